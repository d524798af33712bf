"""Test-only CPU oracle for the GPT-ST pre-training hot path (see gptst_oracle.py header)."""

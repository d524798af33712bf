for v in 0 1 0 1; do echo -n "PDL=$v cap-only: "; GPTST_B200_PDL=$v timeout 60 python bench.py --cap-only 2>/dev/null | grep '^{' | cut -c1-200; done
for v in 1 0; do echo -n "PDL=$v step: "; GPTST_B200_PDL=$v timeout 100 python bench.py --no-cpu-baseline --no-rooflines 2>/dev/null | grep '^{' | cut -c1-120; done

#!/bin/bash
# "Install" the UNMODIFIED reference under baseline/_ref (git-ignored, NOT gpurun-ignored: it travels to the GPU box).
# The reference is a tree of scripts (no setup.py / pyproject: `pip install /root/reference` has nothing to build), so the
# install is a plain copy of what the pre-training path and config 5 need: conf/, lib/, model/ (with the shipped checkpoints)
# and the PEMS08 data set, unzipped where lib/load_dataset.py:45 looks for it.  Nothing under baseline/_ref is edited.
#   tools/install_reference.sh [/root/reference]
set -euo pipefail
SRC=${1:-/root/reference}
DST="$(cd "$(dirname "$0")/.." && pwd)/baseline/_ref"
[ -f "$SRC/model/Pretrain_model/GPTST.py" ] || { echo "no reference at $SRC"; exit 1; }
mkdir -p "$DST/data"
cp -r "$SRC/conf" "$SRC/lib" "$SRC/model" "$SRC/readme.md" "$SRC/requirements.txt" "$SRC/LICENSE.txt" "$DST/"
chmod -R u+w "$DST"
if [ ! -f "$DST/data/PEMS08/PEMS08.npz" ]; then
    python - "$SRC/data/PEMS08.zip" "$DST/data" <<'PY'
import sys, zipfile
zipfile.ZipFile(sys.argv[1]).extractall(sys.argv[2])
PY
fi
find "$DST" -name __pycache__ -type d -prune -exec rm -rf {} +
du -sh "$DST"; ls "$DST/data/PEMS08"

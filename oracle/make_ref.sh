#!/bin/bash
# Copies the UNMODIFIED reference modules of the hot path into the git-ignored oracle/_ref/ so that they travel to
# the GPU box with the repository snapshot (the box has no /root/reference).  Nothing here is committed: oracle/_ref/
# is listed in .gitignore.  Used only by bench.py's reference arms (`--impl reference`, `cpu_baseline`,
# `gpu_eager_baseline`) and by tests that compare against the live reference.
#   bash oracle/make_ref.sh [/root/reference]
set -e
SRC=${1:-/root/reference}
DST="$(cd "$(dirname "$0")" && pwd)/_ref"
if [ ! -f "$SRC/models.py" ]; then
  echo "make_ref: $SRC not found (GPU box?) -- keeping whatever is in $DST"
  exit 0
fi
rm -rf "$DST"
mkdir -p "$DST/configs" "$DST/data"
cp "$SRC/models.py" "$SRC/loss_functions.py" "$SRC/utils.py" "$SRC/train.py" "$SRC/eval.py" "$DST/"
cp "$SRC/configs/config.py" "$SRC/configs/config_ex.py" "$DST/configs/"
cp "$SRC/data/NeuronDataset.py" "$DST/data/"
echo "make_ref: copied the reference's hot-path modules to $DST"

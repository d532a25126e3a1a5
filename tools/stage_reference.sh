#!/bin/bash
# Stage a verbatim, git-ignored copy of the reference's operator files under baseline/_ref/ 
# so that the GPU box -- which has no /root/reference -- can run the reference Triton kernels beside ours.  Only
# tools/triton_parity.py, tests marked "needs baseline/_ref" and bench-side baseline scripts import from there; product
# code never does.  baseline/_ref/ is listed in .gitignore (never committed) and NOT in .gpurunignore (travels with gpurun).
set -e
ROOT=$(cd "$(dirname "$0")/.." && pwd)
REF=${1:-/root/reference}
DST=$ROOT/baseline/_ref
[ -d "$REF/src/model/ops" ] || { echo "reference not found at $REF"; exit 1; }
mkdir -p $DST/src/model/ops $DST/src/utils
cp $REF/src/model/ops/flash_attention_v2_bias.py $REF/src/model/ops/rms_norm.py $REF/src/model/ops/cross_entropy_loss.py $DST/src/model/ops/
cp $REF/src/utils/attn_ref.py $DST/src/utils/
( cd $REF && git rev-parse HEAD 2>/dev/null || cat .SUBMODULES.json 2>/dev/null | head -5 ) > $DST/REVISION.txt 2>/dev/null || true
ls -la $DST/src/model/ops $DST/src/utils

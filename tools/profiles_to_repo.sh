#!/bin/bash
# gpurun_out/<TAG>_* (ncu scratch) -> profiles/<TAG>_* (tracked summaries).  Usage: tools/profiles_to_repo.sh r2b
cd "$(dirname "$0")/.."
TAG=${1:-r2b}
cp gpurun_out/${TAG}_launches.csv profiles/${TAG}_launches.csv
python tools/summarize_profiles.py launches gpurun_out/${TAG}_launches.csv profiles/${TAG}_launches_summary.md
declare -A NOTE=(
 [l1_conv2]="L1 3x3 conv2 fprop 256->192 (fp32 out)."
 [c3x3_conv2]="L0 3x3 conv2 fprop 256->48 (fp32 out): small-N MMAs."
 [c3x3_conv1]="L0 3x3 conv1 fprop 24->256 + ReLU (bf16 out, 67 MB)."
 [wgg_l1_3x3]="Grouped weight + bias gradients of one level-1 3x3 coupling block (4 problems, one launch)."
 [wgg_l0_3x3]="Grouped weight + bias gradients of one level-0 3x3 coupling block (4 problems, one launch)."
 [bwd1x1]="Fused backward of a level-0 1x1 subnet (subnet1x1_bwd.cu): hidden re-evaluated on chip, input gradient, both weight / bias gradients."
 [s1x1_fwd]="Fused level-0 1x1 subnet forward, nothing kept."
 [coupling_bwd]="Standalone coupling backward at level 0 (7T * L/C algorithmic bytes = 88 MB)."
)
for c in "${!NOTE[@]}"; do
  [ -f gpurun_out/${TAG}_${c}.ncu-rep ] && python tools/summarize_profiles.py rep gpurun_out/${TAG}_${c}.ncu-rep profiles/${TAG}_${c}.md "${NOTE[$c]}"
done

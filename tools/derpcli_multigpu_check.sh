#!/bin/bash
# DerpCLI --gpus=2 must write byte-identical PFMs to --gpus=1 (frames are sharded, nothing else changes)
set -e
T=$(mktemp -d)
python - "$T" <<'PY'
import sys, os
sys.path.insert(0, os.getcwd())
from facebook360_dep_b200 import synth
from tests.test_apps import write_dataset
rig = synth.ring_rig(5, 96, 96, kind="FTHETA")
frames = [synth.render_rig(rig, 96, 96, scene=synth.Scene(seed=42, shift=(0.01 * f, 0, 0)))[0] for f in range(4)]
write_dataset(sys.argv[1] + "/in", rig, frames, 2)
PY
B=facebook360_dep_b200/bin/DerpCLI
$B --input_root=$T/in --output_root=$T/out1 --first=000000 --last=000003 --partial_coverage --num_depths=32 --gpus=1 2>/dev/null
$B --input_root=$T/in --output_root=$T/out2 --first=000000 --last=000003 --partial_coverage --num_depths=32 --gpus=2 2>/dev/null
# one frame on two GPUs: destination cameras are sharded instead
$B --input_root=$T/in --output_root=$T/out3 --first=000001 --last=000001 --partial_coverage --num_depths=32 --gpus=2 2>/dev/null
for f in $(cd $T/out3 && find . -name "*.pfm"); do cmp $T/out3/$f $T/out1/$f; done && echo "camera-sharded frame identical: $(find $T/out3 -name "*.pfm" | wc -l) PFMs"
# the same with mismatch handling on the fine level: the GPUs exchange their disparity planes before the stage
$B --input_root=$T/in --output_root=$T/out4 --first=000001 --last=000001 --partial_coverage --num_depths=32 --mismatches_start_level=0 --gpus=1 2>/dev/null
$B --input_root=$T/in --output_root=$T/out5 --first=000001 --last=000001 --partial_coverage --num_depths=32 --mismatches_start_level=0 --gpus=2 2>/dev/null
diff -r $T/out4 $T/out5 && echo "camera-sharded frame with mismatch handling identical: $(find $T/out5 -name '*.pfm' | wc -l) PFMs"
if diff -rq $T/out4/disparity_levels/level_0 $T/out1/disparity_levels/level_0 >/dev/null 2>&1; then echo "note: mismatch handling changed nothing on this data"; fi
# temporal filter over the 4 frames of out1's level-0 disparities: frame blocks on two GPUs vs one
TB=facebook360_dep_b200/bin/TemporalBilateralFilter
cp -r $T/out1 $T/out1b
$TB --input_root=$T/in --output_root=$T/out1 --rig=$T/in/rigs/rig_calibrated.json --first=000000 --last=000003 --level=0 --gpus=1 2>/dev/null
$TB --input_root=$T/in --output_root=$T/out1b --rig=$T/in/rigs/rig_calibrated.json --first=000000 --last=000003 --level=0 --gpus=2 2>/dev/null
diff -r $T/out1/disparity_time_filtered_levels $T/out1b/disparity_time_filtered_levels && echo "TemporalBilateralFilter --gpus=2 == --gpus=1 : $(find $T/out1b/disparity_time_filtered_levels -name '*.pfm' | wc -l) PFMs identical"
rm -rf $T/out1/disparity_time_filtered_levels
diff -r $T/out1 $T/out2 && echo "DerpCLI --gpus=2 == --gpus=1 : $(find $T/out2 -name '*.pfm' | wc -l) PFMs identical"
rm -rf $T

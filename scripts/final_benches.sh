#!/bin/bash
# Round-end bench lines on one B200: the headline config plus BASELINE.json configs[0..3] at their own batch sizes (eager and,
# where launch-bound, replayed from a CUDA graph), the other fused engines, and the 100x100 / latent-512 architecture block.
OUT=gpurun_out/final
mkdir -p $OUT
python bench.py --steps 20 --warmup 5 > $OUT/stage1_B4096_1gpu.json 2> $OUT/stage1_B4096_1gpu.err
python bench.py --impl reference --steps 3 --warmup 1 > $OUT/stage1_reference_arm_cpu.json 2> /dev/null
X="--steps 20 --warmup 5 --no-stock-torch"
python bench.py --batch 64 $X > $OUT/stage1_B64_1gpu.json 2> /dev/null
python bench.py --batch 64 --graph $X --no-cpu-baseline > $OUT/stage1_B64_1gpu_cudagraph.json 2> /dev/null
python bench.py --workload stage1_waegan --batch 256 $X > $OUT/stage1_waegan_B256_1gpu.json 2> /dev/null
python bench.py --workload stage1_waegan --batch 256 --graph $X --no-cpu-baseline > $OUT/stage1_waegan_B256_1gpu_cudagraph.json 2> /dev/null
python bench.py --workload stage1_wae_mmd --batch 256 $X --no-cpu-baseline > $OUT/stage1_wae_mmd_B256_1gpu.json 2> /dev/null
python bench.py --workload stage2_cognitive --batch 256 $X > $OUT/stage2_cognitive_B256_1gpu.json 2> /dev/null
python bench.py --workload stage2_cognitive --batch 256 --graph $X --no-cpu-baseline > $OUT/stage2_cognitive_B256_1gpu_cudagraph.json 2> /dev/null
python bench.py --workload stage3_dual --batch 512 $X > $OUT/stage3_dual_B512_1gpu.json 2> /dev/null
python bench.py --workload stage3_cognitive --batch 512 $X --no-cpu-baseline > $OUT/stage3_cognitive_B512_1gpu.json 2> /dev/null
python bench.py --workload stage2_wae_cognitive --batch 256 $X --no-cpu-baseline > $OUT/stage2_wae_cognitive_B256_1gpu.json 2> /dev/null
python bench.py --workload stage3_wae_cognitive --batch 256 $X --no-cpu-baseline > $OUT/stage3_wae_cognitive_B256_1gpu.json 2> /dev/null
python bench.py --resolution 100 --batch 2048 --steps 10 --warmup 3 > $OUT/stage1_res100_B2048_1gpu.json 2> /dev/null
python scripts/step_profile.py > $OUT/per_call_B4096.txt 2>&1
for f in $OUT/*.json; do echo "$(basename $f): $(python -c "import json,sys; d=json.loads([l for l in open('$f') if l.startswith('{')][-1]); print(round(d['value'],1), d['unit'], round(d['ms_per_step'],3), 'ms')" 2>&1)"; done

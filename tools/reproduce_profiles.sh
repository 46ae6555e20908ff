#!/usr/bin/env bash
# Regenerates the measured evidence under profiles/ on a B200 box (run from the repo root, library built:
# `python __graft_entry__.py`).  Every ncu pass runs only after the same command has exited 0 without ncu;
# numbers printed under ncu are never bench values.  Output goes to $OUT (default gpurun_out/); copy what
# should be kept into profiles/ under the tag of the round (see profiles/README.md).
set -euo pipefail
OUT=${OUT:-gpurun_out}; TAG=${TAG:-rXX}
mkdir -p "$OUT"

# 1. headline bench, reference arm, multi-GPU (N = 2, 4, 8 when the box has them)
python bench.py > "$OUT/${TAG}_bench.json"
python bench.py --impl reference --steps 3 --warmup 1 > "$OUT/${TAG}_bench_reference_arm.json"
# python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 8 --steps 10 --warmup 3 > "$OUT/${TAG}_bench_8gpu.json"

# 2. ncu: launch list of the default bench command, then the full capture of one step's three batch kernels
#    (k_batch_pyramid + one k_batch_level per active level; under ncu the overlapped level launches are serialised)
python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-secondary > /dev/null
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:k_batch --csv --log-file "$OUT/${TAG}_launches_default_bench.csv" \
    python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-secondary > /dev/null
python bench.py --pairs 2368 --steps 2 --warmup 3 --no-cpu-baseline --no-secondary > "$OUT/prof_${TAG}_bench.log"
ncu --set full --clock-control none --import-source on -k regex:k_batch --launch-skip 9 -c 3 -f -o "$OUT/prof_${TAG}" \
    python bench.py --pairs 2368 --steps 2 --warmup 3 --no-cpu-baseline --no-secondary > /dev/null
python tools/ncu_summary.py "$OUT/prof_${TAG}.ncu-rep" > "$OUT/${TAG}_k_batch_ncu_full.txt"
python tools/make_roofline_inputs.py "$OUT/prof_${TAG}.ncu-rep" "$OUT/prof_${TAG}_bench.log" profiles/r01_fp64_peak.jsonl "$TAG" > "$OUT/roofline_inputs.json"

# 3. parity sweeps (GPU vs the CPU oracle)
python tools/fuzz_parity.py --groups 1466 --pairs 24 --seed 1 > "$OUT/${TAG}_fuzz_parity.json"
python tools/fullsize_parity.py --pairs 4096 > "$OUT/${TAG}_fullsize_parity_4096_pairs.json"
python tools/extreme_sizes_check.py > "$OUT/${TAG}_extreme_cases.txt"

# 4. latency of the per-pair API (BASELINE configs 0, 1, 2, 4) and the recorded-sequence reader
python tools/bench_latency.py --mode single > "$OUT/${TAG}_latency_single_pair.json"
python tools/bench_latency.py --mode single --path 3 > "$OUT/${TAG}_latency_single_pair_cluster_driver.json"
python tools/bench_latency.py --mode vo --frames 1000 > "$OUT/${TAG}_latency_vo_sequence_1000frames.json"
python tools/bench_latency.py --mode ceres > "$OUT/${TAG}_latency_ceres_config.json"
python bench.py --workload 8k > "$OUT/${TAG}_bench_8k_1gpu.json"     # N > 1: torchrun ... bench.py --workload 8k --gpus N
python tools/run_row_sharded.py > "$OUT/${TAG}_row_sharded_8k_1gpu.json"
python tools/bench_dataset.py > "$OUT/${TAG}_dataset_vo_on_disk.json"

# 5. hardware probes the rooflines lean on
for p in fp64_peak fp64_latency dmma_probe issue_probe; do
  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o "tools/$p" "tools/$p.cu" && "./tools/$p" > "$OUT/${TAG}_$p.jsonl"
done

# 6. the wave path of the batch entry (Ceres-mode, photometric + depth solver, blurred / large levels)
python tools/bench_pool.py 4 2048 > "$OUT/${TAG}_wave_path_throughput.json"            # pairs/s of both other solvers
python tools/fuzz_parity.py --wave --groups 1500 --pairs 12 --seed 7 > "$OUT/${TAG}_fuzz_parity_wave_path_18k_pairs.json"
ncu --set full --clock-control none --import-source on -k regex:k_align_slots --launch-skip 2 -c 1 -f -o "$OUT/prof_${TAG}_slots_bi" \
    python tools/bench_pool.py 4 592 > /dev/null && python tools/ncu_summary.py "$OUT/prof_${TAG}_slots_bi.ncu-rep" > "$OUT/${TAG}_k_align_slots_bi.txt"
ncu --set full --clock-control none --import-source on -k regex:k_align_slots --launch-skip 6 -c 1 -f -o "$OUT/prof_${TAG}_slots_ceres" \
    python tools/bench_pool.py 4 592 > /dev/null && python tools/ncu_summary.py "$OUT/prof_${TAG}_slots_ceres.ncu-rep" > "$OUT/${TAG}_k_align_slots_ceres.txt"

# 7. where the clocks of k_batch_level go (a variant library with clock64() around the sections of an iteration)
bash tools/build_variant.sh secclk -- -DPHOVO_SECTION_CLOCKS && python tools/section_clocks.py build/variants/secclk.so 4096 > "$OUT/${TAG}_section_clocks.json"

# 8. pool vs slot waves vs clusters by batch size; randomised sweeps judged by the CPU oracle AND the reference's own headers (oracle/_ref)
python tools/wave_cluster_sweep.py 4 8 12 16 24 32 48 64 96 128 148 200 296 400 592 1184 > "$OUT/${TAG}_wave_cluster_sweep.jsonl"
python tools/fuzz_parity.py --groups 1466 --pairs 24 --seed 1 > "$OUT/${TAG}_fuzz_parity_35k_pairs.json"
python tools/fuzz_parity.py --groups 500 --pairs 24 --seed 5 > "$OUT/${TAG}_fuzz_parity_12k_pairs.json"
for pp in "40 80" "100 40" "310 12"; do set -- $pp; python tools/fuzz_parity.py --wave --groups $2 --pairs $1 --seed 31 | tail -1; done > "$OUT/${TAG}_fuzz_parity_wave_path_cluster_sizes.jsonl"
bash tools/build_variant.sh cminb2 -- -DPHOVO_CERES_COOP_MINB=2 && bash tools/ceres_variants.sh     # k_level_coop_ceres at 2 vs 1 CTAs per SM

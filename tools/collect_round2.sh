#!/bin/bash
# Collects what profiles/r02/ is built from (one GPU): tests, bench lines, ncu launch list of the bench command,
# ncu --set full of the hot kernels (raw + details pages as CSV), crossover tables, the config sweep.
# usage (under gpurun): tools/collect_round2.sh
out=gpurun_out
tag=r02
python -m pytest tests -m gpu -q 2>&1 | tail -3 > $out/${tag}_pytest.txt
python bench.py > $out/${tag}_bench_n1.json 2> $out/${tag}_bench_n1.err
python bench.py --impl reference --steps 2 --warmup 1 > $out/${tag}_bench_ref.json 2> $out/${tag}_bench_ref.err
B="python bench.py --steps 2 --warmup 1 --no_e2e --no_cpu_baseline --no_configs --no_policy"
$B > $out/${tag}_bench_short.json 2> /dev/null &&
  ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file $out/${tag}_launches.csv $B > $out/${tag}_ncu_launches.log 2>&1
tools/ncu_capture.sh ${tag}_cov_step coverage_step_kernel 60 $B
tools/ncu_capture.sh ${tag}_cov_returns returns_kernel 1 $B
P="python tools/profile_env.py coverage 64 32 1048576 6"
$P > /dev/null && tools/ncu_capture.sh ${tag}_cov_coop_step_a32 coverage_coop_step 3 $P
P="python tools/profile_env.py congestion 10 8 1048576 20"
$P > /dev/null && tools/ncu_capture.sh ${tag}_cong_step_a8 congestion_step_kernel 5 $P && tools/ncu_capture.sh ${tag}_cong_roll_a8 congestion_rollout 0 $P
P="python tools/profile_env.py congestion 64 32 1048576 6"
$P > /dev/null && tools/ncu_capture.sh ${tag}_cong_coop_step_a32 congestion_coop_step 3 $P && tools/ncu_capture.sh ${tag}_cong_coop_roll_a32 congestion_coop_rollout 0 $P
P="python tools/profile_env.py congestion 32 16 1048576 6"
$P > /dev/null && tools/ncu_capture.sh ${tag}_cong_coop_step_a16 congestion_coop_step 3 $P
P="python tools/profile_env.py collision 64 32 1048576 6"
$P > /dev/null && tools/ncu_capture.sh ${tag}_coll_coop_step_a32 collision_coop_step 3 $P && tools/ncu_capture.sh ${tag}_coll_coop_roll_a32 collision_coop_rollout 0 $P
P="python tools/profile_env.py collision 32 16 1048576 6"
$P > /dev/null && tools/ncu_capture.sh ${tag}_coll_coop_step_a16 collision_coop_step 3 $P
P="python tools/profile_env.py collision 5 3 1048576 20"
$P > /dev/null && tools/ncu_capture.sh ${tag}_coll_step_a3 collision_step_kernel 5 $P
P="python tools/profile_policy.py 16 1048576 2"      # tensor-core build (default), two tiles per thread
$P > /dev/null && tools/ncu_capture.sh ${tag}_policy_tc_a16 policy_act_discrete_tc 3 $P
P="python tools/profile_policy.py 16 262144 0"       # FP32-pipe build
$P > /dev/null && tools/ncu_capture.sh ${tag}_policy_a16 policy_act_discrete_kernel 3 $P
for a in 3 8 16 32; do python tools/profile_policy.py $a 1048576 --time; done > $out/${tag}_policy_variants_time.txt 2>&1
for c in "3 65536" "3 1048576" "8 1048576" "16 262144"; do python tools/time_policy_gauss.py $c; done > $out/${tag}_policy_gauss_time.txt 2>&1
timeout 60 tools/umma_probe > $out/${tag}_umma_probe.txt 2>&1
timeout 60 tools/fma_probe > $out/${tag}_fma_probe.txt 2>&1
for f in $out/${tag}_*_source.csv; do python tools/ncu_summary.py ${f%_source.csv} > ${f%_source.csv}_summary.txt 2>&1; done
rm -f $out/${tag}_*_source.csv            # large; the raw + details pages and the summaries are what profiles/ keeps
python tools/time_coop.py --envs collision,congestion,coverage --agents 12,16,20,24,28,32 > $out/${tag}_crossover.md 2>&1
python tools/time_coop.py --envs congestion --agents 9,12,16,20,24,28,32 --lanes 0,4 > $out/${tag}_crossover_congestion_rollout.md 2>&1
python tools/sweep.py > $out/${tag}_sweep.md 2> $out/${tag}_sweep.err
# programmatic dependent launch on / off on the launch-bound shapes (closed loop as one CUDA graph)
CF="--cfg coverage,5,3,50,50 --cfg collision,5,3,65536,50 --cfg coverage,5,3,65536,50 --cfg congestion,10,8,65536,100 --cfg coverage,32,16,524288,50 --cfg congestion,10,8,1048576,100"
for v in 0 1 0 1; do echo "## SMARL_PDL=$v"; SMARL_PDL=$v python tools/sweep.py $CF | tail -6; done > $out/${tag}_pdl_on_off.md 2>&1
python examples/train_coverage.py > $out/${tag}_example.txt 2>&1
tail -2 $out/${tag}_pytest.txt; tail -c 300 $out/${tag}_bench_n1.json; ls $out | grep ${tag}_ | wc -l

#!/bin/bash
# Collects the measurements a round's profiles/ directory is built from (one GPU):
#   tests, bench line, ncu launch list of the bench command, ncu --set full of every hot kernel, config sweep.
# usage (under gpurun): tools/collect_round.sh <tag>
tag=${1:-final}
out=gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -3 > $out/${tag}_pytest.txt
python bench.py > $out/${tag}_bench_n1.json 2> $out/${tag}_bench_n1.err
python bench.py --impl reference --steps 2 --warmup 1 > $out/${tag}_bench_ref.json 2> $out/${tag}_bench_ref.err
B="python bench.py --steps 2 --warmup 1 --no_e2e --no_cpu_baseline"
$B > $out/${tag}_bench_short.json 2> /dev/null &&
  ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $out/${tag}_launches.csv $B > $out/${tag}_ncu_launches.log 2>&1
tools/ncu_capture.sh ${tag}_cov_step coverage_step 60 $B
tools/ncu_capture.sh ${tag}_cov_returns returns_kernel 1 $B
tools/ncu_capture.sh ${tag}_cov_roll coverage_rollout 1 $B
P="python tools/profile_env.py congestion 10 8 1048576 20"
$P > /dev/null && tools/ncu_capture.sh ${tag}_cong_step congestion_step 5 $P && tools/ncu_capture.sh ${tag}_cong_roll congestion_rollout 0 $P
P="python tools/profile_env.py collision 5 3 1048576 20"
$P > /dev/null && tools/ncu_capture.sh ${tag}_coll_step collision_step 5 $P && tools/ncu_capture.sh ${tag}_coll_roll collision_rollout 0 $P
rm -f $out/${tag}_*_source.csv            # large; the raw + details pages are what profiles/ keeps
python tools/sweep.py > $out/${tag}_sweep.md 2> $out/${tag}_sweep.err
python examples/train_coverage.py > $out/${tag}_example.txt 2>&1
tail -2 $out/${tag}_pytest.txt; tail -c 300 $out/${tag}_bench_n1.json; ls $out | grep ${tag}_ | wc -l

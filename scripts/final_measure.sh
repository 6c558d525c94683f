#!/bin/bash
# Final measurement pass of a session (run under gpurun, one GPU): bench lines of every workload, the reference arm,
# PPO lines, the ncu launch list of the default bench command and --set full captures of K1 and the ObjLock step.
set -u
O=gpurun_out
T=${1:-s2}
python bench.py --steps 2000 --warmup 100 > $O/${T}_bench.json 2> $O/${T}_bench.err || exit 1
python bench.py --impl reference --steps 3 --warmup 1 > $O/${T}_ref.json 2>> $O/${T}_bench.err
for w in waypoints_v3 waypoint_objlock lowlevel objlock_duck; do
  python bench.py --workload $w --steps 1000 --warmup 100 --no-cpu-baseline > $O/${T}_$w.json 2>> $O/${T}_bench.err
done
python bench.py --steps 1000 --warmup 50 --steps-per-launch 8 --no-cpu-baseline > $O/${T}_fused8.json 2>> $O/${T}_bench.err
python bench.py --steps 300 --warmup 20 --envs 1048576 --no-cpu-baseline > $O/${T}_1m.json 2>> $O/${T}_bench.err
python bench.py --workload ppo --steps 5 --warmup 2 > $O/${T}_ppo4096.json 2>> $O/${T}_bench.err
python bench.py --workload ppo --steps 5 --warmup 2 --ppo-envs 65536 --ppo-n-steps 64 > $O/${T}_ppo65536.json 2>> $O/${T}_bench.err
python bench.py --workload ppo --steps 5 --warmup 2 --ppo-envs 65536 --ppo-n-steps 64 --ppo-preset waypoint_objlock > $O/${T}_ppo_ol.json 2>> $O/${T}_bench.err
python bench.py --workload ppo --steps 5 --warmup 2 --ppo-envs 4096 --ppo-n-steps 128 --ppo-preset objlock_duck > $O/${T}_ppo_duck.json 2>> $O/${T}_bench.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/${T}_launches_default_bench.csv \
    python bench.py --steps 200 --warmup 20 --no-cpu-baseline > $O/${T}_ncu_launches.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:fw_step_kernel --launch-skip 60 --launch-count 1 -f \
    -o $O/prof_physics_${T} python bench.py --steps 100 --warmup 20 --no-cpu-baseline > $O/${T}_ncu_k1.log 2>&1
python scripts/objlock_ncu.py > /dev/null 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:fw_step_objlock --launch-skip 150 --launch-count 1 -f \
    -o $O/prof_objlock_${T}_steady python scripts/objlock_ncu.py > $O/${T}_ncu_ol.log 2>&1
tail -c 400 $O/${T}_bench.json; echo; tail -c 300 $O/${T}_ref.json; echo; ls -la $O | grep ${T} | wc -l

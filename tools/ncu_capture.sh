#!/usr/bin/env bash
# Captures the profiles the bench line cites (run on a GPU box: `gpurun --timeout 1800 -- bash tools/ncu_capture.sh`):
#   1. the bench command itself, without ncu (must exit 0; its number is the only bench value)
#   2. a launch list of the same command (`--metrics gpu__time_duration.sum --clock-control none`): per-launch times, cold-cache and
#      serialised -- the kernels' SHARE of the step is what must agree with the CUDA-event stage times of (1)
#   3. one `--set full` capture of the solver kernels of one tick of the aged world (source pages included)
# and writes gpurun_out/<tag>_launches.csv, gpurun_out/<tag>_solver.ncu-rep and profiles/ncu_traffic.json: the DRAM bytes per tick of the
# dominant stages keyed by the hash of the CUDA sources and by the workload, which bench.py prints as roofline.traffic only when both match.
set -u
TAG=${1:-r2}
ORG=${2:-100000}
AGE=${3:-3000}
OUT=gpurun_out
REP=/tmp/sapio_ncu            # the .ncu-rep files stay on the box (a full capture of 16 kernels is ~100 MB; gpurun copies back at most 64 MiB): text summaries travel
mkdir -p $OUT $REP profiles
BENCH="python bench.py --no-cpu --no-forward-record --organisms $ORG --age $AGE --steps 10 --warmup 5"
$BENCH > $OUT/${TAG}_plain.json 2> $OUT/${TAG}_plain.err || { echo "bench failed"; tail -5 $OUT/${TAG}_plain.err; exit 1; }
# solver kernels per tick: 8 size classes + 7 component tiers + the cooperative fallback
PER_TICK=16
SKIP=$(( (AGE + 5) * PER_TICK ))
timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"k_org_solve|k_component|k_coupled" -s $SKIP -c $(( 4 * PER_TICK )) --csv \
    --log-file $OUT/${TAG}_launches.csv $BENCH > $OUT/${TAG}_ncu_launch.log 2>&1
timeout 1500 ncu --set full --import-source on --clock-control none -k regex:"k_org_solve|k_component|k_coupled" -s $SKIP -c $PER_TICK -f -o $REP/${TAG}_solver \
    $BENCH > $OUT/${TAG}_ncu_full.log 2>&1
# every kernel of a few ticks of a 600-tick world (the streaming stages do not depend on the age): grid build, narrowphase, rules, think, train
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 30000 -c 300 --csv --log-file $OUT/${TAG}_launches_tick.csv \
    python bench.py --no-cpu --no-forward-record --organisms $ORG --age 600 --steps 10 --warmup 5 > $OUT/${TAG}_ncu_tick.log 2>&1
python tools/ncu_summary.py launches $OUT/${TAG}_launches_tick.csv > profiles/${TAG}_launches_tick.txt 2>&1
python tools/ncu_traffic.py $REP/${TAG}_solver.ncu-rep $ORG $AGE profiles/ncu_traffic.json
python tools/ncu_summary.py launches $OUT/${TAG}_launches.csv > profiles/${TAG}_launches.txt 2>&1
python tools/ncu_summary.py full $REP/${TAG}_solver.ncu-rep > profiles/${TAG}_solver_full.txt 2>&1
tail -3 profiles/ncu_traffic.json
# profiles/ is outside gpurun_out/: bring the summaries back through it (copy them into profiles/ of the working tree afterwards)
mkdir -p $OUT/profiles && cp profiles/ncu_traffic.json profiles/${TAG}_launches.txt profiles/${TAG}_launches_tick.txt profiles/${TAG}_solver_full.txt $OUT/profiles/ 2>/dev/null

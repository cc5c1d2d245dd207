#!/bin/bash
# Run on the GPU box (gpurun -- bash tools/capture_profiles.sh <round tag>): one `ncu --set full` capture of k_trace per
# workload (C2, C3, C4 and the 1e7-triangle sweep point), folded into profiles/traffic.json by tools/ncu_traffic.py, plus
# the per-kernel summaries and the launch list of the bench command.  Every program first runs once WITHOUT ncu.
# Only the C2 report keeps its .ncu-rep (source page); gpurun_out/ is limited to 64 MiB.
set -u
TAG=${1:-r02}
OUT=gpurun_out
mkdir -p $OUT
capture() {   # key, tris override ("" = none), workload
    local key=$1 tris=$2 wl=$3
    local extra=""
    [ -n "$tris" ] && extra="--tris $tris"
    python tools/prof_scan.py --workload $wl $extra --reps 2 > $OUT/${TAG}_plain_$key.log 2>&1 || { echo "plain run failed for $key"; return 1; }
    ncu --set full --metrics lts__t_bytes.sum,l1tex__t_bytes.sum --clock-control none --import-source on -k regex:k_trace \
        --launch-skip 1 --launch-count 1 -f -o $OUT/prof_${TAG}_$key python tools/prof_scan.py --workload $wl $extra --reps 2 > $OUT/${TAG}_ncu_$key.log 2>&1
    local line
    line=$(tail -1 $OUT/${TAG}_plain_$key.log)
    local T R B
    T=$(python -c "import ast,sys; print(ast.literal_eval(sys.argv[1])['tris'])" "$line")
    R=$(python -c "import ast,sys; print(ast.literal_eval(sys.argv[1])['rays_per_launch'])" "$line")
    B=$(python -c "import ast,sys; print(ast.literal_eval(sys.argv[1])['build_tag'])" "$line")
    python tools/ncu_traffic.py $OUT/prof_${TAG}_$key.ncu-rep $key $T $R $B \
        "ncu --set full --clock-control none, second k_trace launch of \`python tools/prof_scan.py --workload $wl $extra --reps 2\` (L2 flushed before the launch); summary in profiles/${TAG}_k_trace_$key.md" > $OUT/${TAG}_traffic_$key.json
    python tools/ncu_summary.py kernel $OUT/prof_${TAG}_$key.ncu-rep $OUT/${TAG}_k_trace_$key.md "ncu --set full of k_trace, workload $key ($wl $extra)" > /dev/null
    [ "$key" != "c2" ] && rm -f $OUT/prof_${TAG}_$key.ncu-rep
    echo "captured $key"
}
capture c2 "" c2
capture c3 "" c3
capture c4 "" c4
capture sweep_1e7 10000000 c2
cp profiles/traffic.json $OUT/${TAG}_traffic.json
# launch list of the bench command (after the same command has run without ncu)
python bench.py --steps 2 --warmup 3 --no-cpu --no-per-frame --extra none > $OUT/${TAG}_bench_plain.json 2> $OUT/${TAG}_bench_plain.err && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/${TAG}_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu --no-per-frame --extra none > $OUT/${TAG}_launches_run.log 2>&1
echo done

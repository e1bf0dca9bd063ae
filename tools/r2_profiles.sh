# Round-2 ncu evidence, one call on one B200 (run through gpurun; the summaries are made from the
# .ncu-rep files afterwards with tools/ncu_summary.py, tools/launch_summary.py, tools/traffic_from_ncu.py).
set -x
OUT=gpurun_out
B="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --e2e-steps 1 --fastq-reads 0 --configs c3"
# every command first runs WITHOUT ncu and must exit 0
$B > $OUT/r2_prof_bench_plain.json 2> $OUT/r2_prof_bench_plain.err || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $OUT/r2_launches.csv $B > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:count_stream_kernel -s 3 -c 2 -o $OUT/r2_stream $B > /dev/null 2>&1
python tools/wide_kernel_time.py > $OUT/r2_prof_wide_plain.log 2>&1 || exit 1
for k in 16 24 28 30; do
  WIDE_KS=$k ncu --set full --clock-control none -k regex:count_stream_kernel -s 4 -c 1 -o $OUT/r2_family_k$k python tools/wide_kernel_time.py > /dev/null 2>&1
done
python tools/generic_kernel_time.py > $OUT/r2_prof_generic_plain.log 2>&1 || exit 1
ncu --set full --clock-control none --import-source on -k regex:count_lines_kernel -s 2 -c 1 -o $OUT/r2_lines python tools/generic_kernel_time.py > /dev/null 2>&1
DINF_SKIP_CLI=1 python tools/device_inflate_time.py 16777216 > $OUT/r2_prof_dinf_plain.log 2>&1 || exit 1
DINF_SKIP_CLI=1 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file $OUT/r2_dinf_launches.csv python tools/device_inflate_time.py 16777216 > /dev/null 2>&1
DINF_SKIP_CLI=1 ncu --set full --clock-control none --import-source on -k regex:inflate_blocks -c 1 -o $OUT/r2_inflate python tools/device_inflate_time.py 16777216 > /dev/null 2>&1
DINF_SKIP_CLI=1 ncu --set full --clock-control none -k regex:span_extract -c 1 -o $OUT/r2_span_extract python tools/device_inflate_time.py 16777216 > /dev/null 2>&1
ls -la $OUT/r2_*.ncu-rep

#!/bin/bash
# ncu evidence of round 2 (run under gpurun on one B200): launch list of the benchmark step, full captures of the marching kernels and of
# the iteration tail.  Every ncu pass follows a plain run of the same command that exited 0.
set -u
O=gpurun_out
mkdir -p $O
B="python bench.py --steps 20 --warmup 3 --no-strong --no-general --no-cpu-baseline"
$B > $O/r02_prof_plain_bench.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/r02_c2_launches.csv $B > $O/r02_prof_ncu_bench.log 2>&1
$B > /dev/null 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:"sweep_march|iteration_tail" -s 8 -c 2 -o $O/r02_c2_march_tail $B > $O/r02_prof_ncu_full.log 2>&1
P5="python tools/kernel_probe.py ncu c5s"
$P5 > /dev/null 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:"sweep_march" -s 6 -c 1 -o $O/r02_c5s_march_persistent $P5 > $O/r02_prof_ncu_c5s.log 2>&1
P3="python tools/kernel_probe.py ncu c3s"
$P3 > /dev/null 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:"sweep_march" -s 6 -c 1 -o $O/r02_c3s_march_tets $P3 > $O/r02_prof_ncu_c3s.log 2>&1
PK="python -c \"import sys; sys.path.insert(0, 'dg-solver-for-pbte-with-mfem_b200'); import pbte_b200 as pb; print(pb.measure_peaks(0))\""
eval $PK > $O/r02_peaks_plain.log 2>&1 &&
eval ncu --set full --clock-control none -k regex:"dmma_peak|dfma_peak|copy_kernel" -c 6 -o $O/r02_peak_probes $PK > $O/r02_prof_ncu_peaks.log 2>&1
[ -f $O/r02_peak_probes.ncu-rep ] && ncu -i $O/r02_peak_probes.ncu-rep --page raw --csv > $O/r02_peak_probes_ncu_raw.csv 2>/dev/null
rm -f $O/r02_peak_probes.ncu-rep
# the raw pages travel back as CSV (gpurun merges at most 64 MiB); only the C2 report itself is kept
for r in r02_c2_march_tail r02_c5s_march_persistent r02_c3s_march_tets; do
    [ -f $O/$r.ncu-rep ] && ncu -i $O/$r.ncu-rep --page raw --csv > $O/${r}_ncu_raw.csv 2>/dev/null
done
ncu -i $O/r02_c2_march_tail.ncu-rep --page source --csv > $O/r02_c2_march_tail_ncu_source.csv 2>/dev/null
rm -f $O/r02_c5s_march_persistent.ncu-rep $O/r02_c3s_march_tets.ncu-rep
ls -la $O

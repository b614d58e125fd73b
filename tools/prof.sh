# usage: bash tools/prof.sh <tag> [kernel-regex] [skip]   -- launch list + one full capture of the hot kernels
# (two ncu passes; on a shared GPU pool run them as two separate gpurun calls: see profiles/README.md)
set -x
TAG=${1:-r1}
KRE=${2:-prep2_kernel|filter_octet|filter_duo}
CMD="python bench.py --frames 8 --steps 1 --warmup 3 --no-cpu-baseline --no-extras"
$CMD > gpurun_out/plain_$TAG.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches_$TAG.csv $CMD > gpurun_out/ncu1_$TAG.log 2>&1
$CMD > gpurun_out/plain2_$TAG.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:"$KRE" -s ${3:-24} -c 2 -o gpurun_out/prof_$TAG $CMD > gpurun_out/ncu2_$TAG.log 2>&1
tail -2 gpurun_out/ncu2_$TAG.log

# A/B of library builds inside one job (clock / box variance between jobs is +-5 %):
#   ACQ_B200_LIB selects the build; usage: bash scripts/ab_variants.sh libA.so libB.so ...
L=$PWD/academicodec_b200/lib
for r in 1 2; do
for v in "$@"; do
  export ACQ_B200_LIB=$L/$v
  echo "== $v"
  timeout 100 python scripts/tc_perf_probe.py 4096 128 100 1024 8
  timeout 100 python scripts/tc_perf_probe.py 64 512 1000 1024 12
  timeout 100 python scripts/tc_perf_probe.py 4096 512 50 1024 2
  timeout 100 python scripts/tc_perf_probe.py
  timeout 100 python scripts/tc_perf_probe.py 16 128 100 1024 8
  timeout 100 python scripts/tc_perf_probe.py 16 512 100 1024 12
done; done

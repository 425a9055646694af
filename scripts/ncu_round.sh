# ncu --set full captures of every kernel on the path at the BASELINE shapes (one GPU; run under gpurun):
#   bash scripts/ncu_round.sh TAG        -> gpurun_out/TAG_*.ncu-rep   (summarise with scripts/ncu_summary.py)
# Every capture follows a plain run of the same command that exited 0 (B200_PROFILING.md).
TAG=${1:-r03}
KEEP=${KEEP:-"search_cfg2 search_cfg4"}
O=gpurun_out
NCU="ncu --set full --clock-control none --import-source on -f"
run() {  # name, kernel regex, skip, command...
  name=$1; regex=$2; skip=$3; shift 3
  "$@" > $O/${TAG}_${name}_plain.log 2>&1 || { echo "plain run failed: $name"; return; }
  $NCU -k regex:$regex -s $skip -c 1 -o $O/${TAG}_${name} "$@" > $O/${TAG}_${name}_ncu.log 2>&1
  tail -1 $O/${TAG}_${name}_ncu.log
  # gpurun brings back at most 64 MiB: keep the raw-metric CSV of every capture, the report itself only for KEEP
  ncu -i $O/${TAG}_${name}.ncu-rep --page raw --csv > $O/${TAG}_${name}_raw.csv 2>/dev/null
  case " $KEEP " in *" $name "*) ;; *) rm -f $O/${TAG}_${name}.ncu-rep ;; esac
}
B="python bench.py --steps 3 --warmup 1 --no-cpu-baseline --no-secondary --no-module --sustained-s 0 --e2e-steps 1"
run search_cfg2 rvq_search 2 python scripts/one_search.py cfg2 0:0 3
run search_cfg1 rvq_search 2 python scripts/one_search.py cfg1 0:0 3
run search_cfg4 rvq_search 2 python scripts/one_search.py cfg4 0:0 3
run search_cfg3 rvq_search 2 python scripts/one_search.py cfg3 0:0 3
run decode_cfg1 vq_decode 3 $B --workload cfg1_b4096
run decode_cfg4 vq_decode 3 $B --workload cfg4_b64
run decode_cfg3 vq_decode 3 $B --workload cfg3_b4096
run decode_cfg2 vq_decode 3 $B --workload cfg2_enc24k_32d_vq1
run replay_cfg5 rvq_replay 3 python scripts/train_step_bench.py 3
run ema_embed_cfg5 ema_embed 3 python scripts/train_step_bench.py 3
run ema_stats "ema_stats" 0 python scripts/ema_stats_probe.py

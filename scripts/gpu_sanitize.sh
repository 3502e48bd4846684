#!/bin/bash
# compute-sanitizer over the small worlds of scripts/sanitize_driver.py: memcheck + racecheck + synccheck on the default path,
# racecheck on the variants.  Usage (under gpurun): bash scripts/gpu_sanitize.sh <tag>
TAG=${1:-san}
OUT=gpurun_out/$TAG
mkdir -p $OUT
CS=/usr/local/cuda/bin/compute-sanitizer
run() {  # name tool env... -- kind steps
  local name=$1 tool=$2; shift 2
  local envs=()
  while [ "$1" != "--" ]; do envs+=("$1"); shift; done
  shift
  ( env "${envs[@]}" timeout 900 $CS --tool $tool --print-limit 20 python scripts/sanitize_driver.py "$@" 2>&1 | tail -40 ) > $OUT/$name.log
  echo "== $name: $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY|sanitize_driver' $OUT/$name.log | tr '\n' ' ')"
}
run memcheck_weibel memcheck -- weibel 3
run racecheck_weibel racecheck -- weibel 2
run synccheck_weibel synccheck -- weibel 2
run racecheck_sm5 racecheck WM_SM=5 -- weibel 2
run memcheck_sm5 memcheck WM_SM=5 -- weibel 3
run memcheck_sm5_slack memcheck WM_SM=5 WM_SLACK=0.4 -- weibel 3
run racecheck_sm3 racecheck WM_SM=3 -- weibel 2
run memcheck_sm3 memcheck WM_SM=3 -- weibel 2
run memcheck_slack memcheck WM_SLACK=0.4 -- weibel 3
run racecheck_slack racecheck WM_SLACK=0.4 -- weibel 2
run memcheck_recon memcheck -- reconnection 3
run racecheck_recon racecheck -- reconnection 2
run memcheck_shock memcheck -- shock 3
run racecheck_shock racecheck -- shock 2
run initcheck_weibel initcheck -- weibel 2
run memcheck_cg0 memcheck WM_CG=0 -- weibel 2
run racecheck_recon_cg racecheck -- reconnection 3

#!/bin/bash
TAG=${1:-r02k}
OUT=gpurun_out
mkdir -p $OUT
for lib in aquaticgymenv_b200/variants/*.so; do
  echo "== $lib"
  AQUAGYM_LIB=$PWD/$lib timeout 600 python tools/many_probe.py quick > $OUT/${TAG}_many_$(basename $lib .so).jsonl 2>> $OUT/${TAG}_many.err
  python - <<PY
import json
for l in open("$OUT/${TAG}_many_$(basename $lib .so).jsonl"):
    j=json.loads(l)
    print("%-24s %-6s n=%-9d K=%-3d v=%d thr=%-3d %s  %.4g  %.3f" % (j["env"], j["mode"], j["envs"], j["K"], j["v"], j["threads"], j["precision"][-2:], j["env_steps_per_s"], j["frac65"]))
PY
done
timeout 600 python tools/replay_shipped.py > $OUT/${TAG}_replay_shipped.jsonl 2> $OUT/${TAG}_replay.err; echo "replay exit $?"; cat $OUT/${TAG}_replay_shipped.jsonl | cut -c1-400

set -x
mkdir -p gpurun_out
for cfg in "1 0" "2 0" "2 132" "2 120" "2 104" "2 88" "1 120"; do
  set -- $cfg
  PADEL_B200_STREAMS=$1 PADEL_B200_BALL_SMS=$2 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r2p_streams$1_sms$2.json 2> gpurun_out/r2p.err || tail -3 gpurun_out/r2p.err
  python - <<PY
import json
d=json.load(open("gpurun_out/r2p_streams$1_sms$2.json"))
print("streams $1 ball_sms $2 :", d["value"], "fps", d["ms_per_step"], "ms  e2e", d["e2e"]["value"])
PY
done

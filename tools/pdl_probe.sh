python -m pytest tests -m gpu -q 2>&1 | tail -6
python tools/small_batch_probe2.py 2>&1 | grep spec
MJXB_PDL=0 python tools/small_batch_probe2.py 2>&1 | grep spec | sed 's/^/nopdl /'
python -c "
import json
from mujoco_mjx_lab_b200 import ppo, apg
print(json.dumps(ppo.time_ppo(1024, 256, iters=5, warmup=3)))
print(json.dumps(apg.time_apg(2048, 128, iters=3, warmup=2)))
" 2>&1 | tail -2

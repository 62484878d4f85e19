#!/usr/bin/env bash
# round 2, GPU session U: generic global-workspace kernel (3_3: N = 8192) with the workspace pinned in L2 (persisting access-policy window)
set +e
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "large_polynomial" > gpurun_out/u_pytest.log 2>&1; tail -2 gpurun_out/u_pytest.log
for p in 1 0; do
  B200_WS_L2_PERSIST=$p timeout 900 python tools/ab_bench.py --lib ours --what set33 --batches 148,296,1024 --steps 3 > gpurun_out/u_set33_p$p.log 2>&1
  echo "persist $p"; grep what gpurun_out/u_set33_p$p.log | python -c "
import sys,json
for l in sys.stdin:
    d=json.loads(l); print('  ',d['batch'],round(d['ms'],3),round(d.get('pbs_per_s',0)))"
done
python - <<'PY'
import torch
p=torch.cuda.get_device_properties(0)
print('L2', p.L2_cache_size, 'persisting max', getattr(p,'persisting_l2_cache_max_size',None), 'window max', getattr(p,'access_policy_max_window_size',None))
PY

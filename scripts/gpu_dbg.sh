#!/bin/bash
mkdir -p gpurun_out
timeout 300 python - <<'PY' 2>&1 | tail -20
import torch as th, sys
sys.path.insert(0,'.')
from morl_baselines_b200 import ops
dev=th.device('cuda:0')
for (M,N,K) in [(64,32,32),(64,6,32),(16,32,32),(130,32,64)]:
    a=th.randn(M,K,device=dev); b=th.randn(N,K,device=dev); bias=th.randn(N,device=dev)
    ap=ops.split_bf16x3(a); bp=ops.split_bf16x3(b, rows_pad=(N+31)//32*32)
    try:
        c,cp=ops.gemm_bf16x3(ap,bp,N,bias=bias,relu=True,out_f32=True,out_planes=(N%32==0))
        th.cuda.synchronize()
        ref=(a.double()@b.double().t()+bias.double()).clamp_min(0)
        print((M,N,K),'ok err',float((c.double()-ref).abs().max()))
    except Exception as e:
        print((M,N,K),'FAILED',repr(e)[:300]); break
PY
timeout 300 python -m pytest tests/test_envelope_update_gpu.py -k api -x --tb=short -q 2>&1 | tail -40
timeout 600 python -m pytest tests/test_gemm_gpu.py -q -x --timeout 300 2>&1 | tail -5
timeout 600 python bench.py --steps 100 --warmup 5 2>&1 | tail -1 | cut -c1-700

mkdir -p gpurun_out
for c in "0 0" "0 1" "1 0" "1 1"; do python scripts/gemm_one.py 8192 8192 8192 $c linear 5; done
python - <<'PY'
import torch
torch.backends.cuda.matmul.allow_tf32 = True
def t(M,N,K,tag=''):
    A=torch.randn(M,K,device='cuda'); B=torch.randn(K,N,device='cuda')
    for _ in range(3): C=A@B
    torch.cuda.synchronize()
    e0,e1=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10): C=A@B
    e1.record(); torch.cuda.synchronize()
    ms=e0.elapsed_time(e1)/10
    print('cublas tf32 %s M=%d N=%d K=%d: %.3f ms %.1f TFLOP/s'%(tag,M,N,K,ms,2*M*N*K/ms/1e9))
t(8192,8192,8192); t(65536,2048,4096); t(65536,4096,2048); t(65536,1024,256)
A=torch.randn(65536,4096,device='cuda'); D=torch.randn(65536,2048,device='cuda')
for _ in range(3): C=A.t()@D
torch.cuda.synchronize()
e0,e1=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10): C=A.t()@D
e1.record(); torch.cuda.synchronize(); ms=e0.elapsed_time(e1)/10
print('cublas tf32 wgrad A^T D (4096x2048, K=65536): %.3f ms %.1f TFLOP/s'%(ms,2*4096*2048*65536/ms/1e9))
a=torch.randn(8192,8192,device='cuda',dtype=torch.bfloat16); b=torch.randn(8192,8192,device='cuda',dtype=torch.bfloat16)
for _ in range(3): c=a@b
torch.cuda.synchronize(); e0.record()
for _ in range(10): c=a@b
e1.record(); torch.cuda.synchronize(); ms=e0.elapsed_time(e1)/10
print('cublas bf16 8192^3: %.3f ms %.1f TFLOP/s'%(ms,2*8192**3/ms/1e9))
PY

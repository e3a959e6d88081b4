import sys, time, torch
sys.path.insert(0, '/root/repo')
import __graft_entry__ as ge; ge.build()
from fish_diffusion_b200 import _native as N
sys.argv=['x']
import importlib.util
spec=importlib.util.spec_from_file_location('bt','/root/repo/tools/bench_train.py'); bt=importlib.util.module_from_spec(spec); spec.loader.exec_module(bt)
from fish_diffusion_b200.train import DenoiserTrainer
dev=torch.device('cuda',0)
diff=bt.build(dev, bt.WN_CFG)
B,T=20,1000
g=torch.Generator().manual_seed(1)
feats=torch.randn(B,T,256,generator=g).to(dev); mel=(torch.rand(B,T,128,generator=g)*5-5).to(dev)
tr=DenoiserTrainer(diff, device=dev)
for _ in range(3): tr.step(feats,mel)
torch.cuda.synchronize()
def ev(): e=torch.cuda.Event(enable_timing=True); e.record(); return e
N.prof_enable(True)
t0=time.perf_counter(); e0=ev()
tr.opt.zero_grad(set_to_none=True)
loss=tr.module(feats,mel); e1=ev(); c1=time.perf_counter()
loss.backward(); e2=ev(); c2=time.perf_counter()
torch.nn.utils.clip_grad_norm_(diff.parameters(),0.5); tr.opt.step(); e3=ev(); c3=time.perf_counter()
torch.cuda.synchronize()
prof,_=N.prof_collect(); N.prof_enable(False)
print('gpu ms: fwd %.2f bwd %.2f opt %.2f'%(e0.elapsed_time(e1), e1.elapsed_time(e2), e2.elapsed_time(e3)))
print('cpu ms: fwd %.2f bwd %.2f opt %.2f'%((c1-t0)*1e3,(c2-c1)*1e3,(c3-c2)*1e3))
print(prof)

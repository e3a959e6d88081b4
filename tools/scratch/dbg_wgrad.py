import ctypes, sys, numpy as np, torch
sys.path.insert(0, '/root/repo')
import __graft_entry__ as ge; ge.build()
from fish_diffusion_b200 import _native as N
dev = torch.device('cuda:0')
B, T, R, Cc = 1, 64, 128, 64
rng = np.random.RandomState(0)
a = torch.from_numpy(rng.randn(B, T, R).astype(np.float32)).to(dev)
c = torch.from_numpy(rng.randn(B, T, Cc).astype(np.float32)).to(dev)
pc = N.PREC_F16
ap, cp = N.split_nwc(a, pc), N.split_nwc(c, pc)
d = N.WgradDesc()
d.row_src[0], d.row_C[0] = N.ptr(ap), R
d.col_src[0], d.col_C[0] = N.ptr(cp), Cc
d.num_row_seg, d.num_col_seg = 1, 1
d.row_seg_src[0], d.row_seg_coff[0], d.row_seg_width[0] = 0, 0, R
d.col_seg_src[0], d.col_seg_shift[0], d.col_seg_coff[0], d.col_seg_width[0] = 0, 0, 0, Cc
part = torch.full((1, R, Cc), float('nan'), dtype=torch.float32, device=dev)
d.B, d.T, d.splits, d.part, d.acc_scale, d.prec = B, T, 1, N.ptr(part), 1.0, pc | N.PREC_SINGLE
rc = N.lib().fd_wgrad_cl(ctypes.byref(d), N.stream_ptr(dev))
print('rc', rc, N.last_error())
torch.cuda.synchronize()
got = part[0].cpu().numpy()
ah = ap[0].cpu().view(torch.float16).double().numpy(); ch = cp[0].cpu().view(torch.float16).double().numpy()
ref = np.einsum('btr,btc->rc', ah, ch)
print('nan count', np.isnan(got).sum(), 'zeros', (got == 0).sum(), 'of', got.size)
print('got[:4,:4]\n', got[:4, :4], '\nref[:4,:4]\n', ref[:4, :4])
print('rel', np.linalg.norm(got - ref) / np.linalg.norm(ref))
# try transposed / other matches
print('rel vs ref.T-ish', np.linalg.norm(got[:64,:64] - ref[:64,:64].T) / np.linalg.norm(ref[:64,:64]))

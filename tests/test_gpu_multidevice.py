"""Modules on a non-current CUDA device: the binding names the device of its tensors (fd_set_device) and
every entry point switches to it for the duration of the call, with per-device kernel attributes.  Needs 2 GPUs."""
import numpy as np
import pytest
import torch

from conftest import rel_l2
from fish_diffusion_b200 import WaveNet
from oracle import wavenet as ownet

pytestmark = pytest.mark.gpu


@pytest.mark.skipif(not torch.cuda.is_available() or torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_module_on_second_device_without_set_device(golden, golden_cfg):
    g = golden("wavenet")
    cfg = golden_cfg["WN_TC"]
    sd = ownet.make_wavenet_weights(12, **{k: v for k, v in cfg.items() if k != "dilation_cycle"})
    assert torch.cuda.current_device() == 0
    outs = []
    for d in (1, 0, 1):                       # second device first: its kernels need their own smem attribute
        dev = torch.device("cuda", d)
        net = WaveNet(**cfg, backend="tc").to(dev)
        net.load_state_dict({k: torch.from_numpy(v) for k, v in sd.items()})
        with torch.no_grad():
            y = net(torch.from_numpy(g["wn_tc_x"]).to(dev), torch.tensor([990], device=dev),
                    torch.from_numpy(g["wn_tc_cond"]).to(dev))
        assert y.device == dev and torch.cuda.current_device() == 0       # the caller's device is restored
        outs.append(y.cpu().numpy())
        assert rel_l2(outs[-1], g["wn_tc_y_t990"]) < 2e-5
    assert np.array_equal(outs[0], outs[2])

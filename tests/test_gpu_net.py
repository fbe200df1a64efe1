"""GPU parity: the policy/value network forward through the C ABI against the golden vectors generated from the
reference's FCModel (fp32, CPU) and against the oracle's restatement.  Tolerance: 1e-5 absolute on logits and values
(BASELINE north_star)."""
import numpy as np
import pytest
import torch

import golden_checks as gc
import gpu_util as gu

pytestmark = pytest.mark.gpu

ATOL = 1e-5


def load_model(sd, S, W, D, tensor_cores=True):
    from boardlaw_b200 import heads
    from boardlaw_b200.networks import FCModel
    net = FCModel(heads.Tensor((S, S, 2)), heads.Masked(S * S), width=W, depth=D)
    net.tensor_cores = tensor_cores
    missing = net.load_state_dict(sd)          # same keys as the reference's state_dict
    assert not missing.missing_keys and not missing.unexpected_keys
    return net.cuda()


def compare(logits, v, ref_logits, ref_v):
    finite = torch.isfinite(ref_logits)
    assert torch.equal(torch.isfinite(logits.cpu()), finite), 'mask differs'
    assert (logits.cpu()[~finite] == ref_logits[~finite]).all()            # -inf on occupied cells
    err_l = (logits.cpu()[finite] - ref_logits[finite]).abs().max().item()
    err_v = (v.cpu() - ref_v).abs().max().item()
    assert err_l <= ATOL, f'logits off by {err_l}'
    assert err_v <= ATOL, f'values off by {err_v}'
    return err_l, err_v


@pytest.mark.parametrize('S,W,D', [(5, 32, 2), (9, 256, 4), (11, 64, 3)])
def test_forward_golden(S, W, D):
    z = np.load(gc.GOLDEN / f'net_S{S}_W{W}_D{D}.npz')
    sd = {k[3:]: torch.from_numpy(z[k]) for k in z.files if k.startswith('sd_')}
    net = load_model(sd, S, W, D)
    logits, v = net.evaluate(torch.from_numpy(z['board']).cuda(), torch.from_numpy(z['seats']).cuda())
    compare(logits, v, torch.from_numpy(z['logits']), torch.from_numpy(z['v']))


@pytest.mark.parametrize('tensor_cores', [True, False])
@pytest.mark.parametrize('S,W,D,B', [(9, 256, 4, 4096), (7, 128, 4, 1000), (13, 64, 2, 517), (3, 2, 4, 64), (11, 512, 8, 512),
                                     (5, 32, 2, 300), (9, 256, 0, 129), (9, 64, 1, 40000),
                                     (11, 512, 8, 20000), (13, 512, 2, 300), (5, 512, 1, 129), (9, 512, 0, 200)])
def test_forward_vs_oracle(S, W, D, B, tensor_cores):
    """tensor_cores=True: tcgen05 split-fp16 kernels where the shape fits (W in 32/64/128/256: net_tc.cu; W = 512:
    net_tc_wide.cu), CUDA-core fp32 kernels otherwise; False forces the CUDA-core path.  Both within 1e-5 of the fp32 CPU
    reference."""
    from oracle import pyref
    sd = pyref.synth_state_dict(S, W, D, seed=S + W)
    w = gu.start_position(S, B, S * S // 2, seed=W)
    ref = pyref.FCNet(sd)(w)
    net = load_model(sd, S, W, D, tensor_cores)
    logits, v = net.evaluate(w.board.cuda(), w.seats.cuda())
    torch.cuda.synchronize()
    el, ev = compare(logits, v, ref.logits, ref.v)
    print(f'S{S} W{W} D{D} tc={tensor_cores}: max |dlogit| {el:.2e}, max |dv| {ev:.2e}')


def test_forward_through_world_api():
    """``network(worlds) -> arrdict(logits, v)`` (boardlaw/networks.py:37-41) on a Hex record, empty batch included."""
    from boardlaw_b200.hex import Hex
    from oracle import pyref
    S, W, D = 5, 32, 2
    sd = pyref.synth_state_dict(S, W, D, seed=0)
    net = load_model(sd, S, W, D)
    w = Hex.initial(10, S, device='cuda')
    d = net(w)
    assert d.logits.shape == (10, 25) and d.v.shape == (10, 2)
    assert torch.allclose(d.logits.exp().sum(-1), torch.ones(10, device='cuda'), atol=1e-5)
    assert torch.equal(d.v[:, 0], -d.v[:, 1])
    e = net(Hex.initial(0, S, device='cuda'))
    assert e.logits.shape == (0, 25)


@pytest.mark.parametrize('S,W,D,B', [(9, 256, 4, 1000), (11, 512, 8, 700), (5, 32, 2, 300)])
def test_forward_amp_mode(S, W, D, B):
    """precision='amp' (fp16 operands, fp32 accumulation — the precision class of the reference's autocast in
    MCTS.simulate, boardlaw/mcts/__init__.py:131-133): one tensor-core product per K-step instead of three.  Within
    half-precision operand error of the fp32 reference, same legal-move mask, and not bit-identical to the fp32 mode."""
    from boardlaw_b200 import heads
    from boardlaw_b200.networks import FCModel
    from oracle import pyref
    sd = pyref.synth_state_dict(S, W, D, seed=S + W)
    w = gu.start_position(S, B, S * S // 2, seed=W)
    ref = pyref.FCNet(sd)(w)
    outs = {}
    for precision in ('fp32', 'amp'):
        net = FCModel(heads.Tensor((S, S, 2)), heads.Masked(S * S), width=W, depth=D, precision=precision)
        net.load_state_dict(sd)
        logits, v = net.cuda().evaluate(w.board.cuda(), w.seats.cuda())
        torch.cuda.synchronize()
        outs[precision] = (logits.cpu(), v.cpu())
    finite = torch.isfinite(ref.logits)
    assert torch.equal(torch.isfinite(outs['amp'][0]), finite)
    err = (outs['amp'][0][finite] - ref.logits[finite]).abs().max().item()
    errv = (outs['amp'][1] - ref.v).abs().max().item()
    print(f'S{S} W{W} D{D} amp: max |dlogit| {err:.2e}, max |dv| {errv:.2e}')
    assert err < 3e-2 and errv < 3e-2
    assert not torch.equal(outs['amp'][0], outs['fp32'][0])

"""The drop-in at op level (INTEGRATION.md level 1): ``boardlaw_b200.cuda.install()`` plants this library's
``hexcuda`` / ``mctscuda`` objects in the reference's two loader caches (boardlaw/hex/cuda.py:4-9, boardlaw/mcts/cuda.py:6-11),
after which the reference's UNCHANGED Python — ``Hex``, ``MCTS``, ``mcts()`` — runs on these kernels.

Where both the reference checkout and a GPU exist, a whole search by the reference's own ``MCTS`` class on the GPU through the
shims is compared, tree for tree, with the same search on the reference's own CPU kernels.  The two never meet on the
round's machines (the reference is absent from the GPU box, the build container has no GPU), so the part that can run
anywhere the reference exists is checked separately: every attribute the reference's Python reads off its native modules
exists on the planted objects, with the call signatures its wrappers use.  (The op-level numerics of the shims are pinned
on the GPU box by the reference-generated recordings: tests/test_gpu_mcts.py::test_mcts_calls_golden, tests/test_gpu_hex.py.)"""
import inspect
import re

import pytest
import torch

from oracle import refpy

needs_ref = pytest.mark.skipif(not refpy.present(), reason='/root/reference is not present on this machine')


@pytest.mark.reference
@needs_ref
def test_planted_objects_cover_what_the_reference_calls():
    from boardlaw_b200.hex import cuda as hex_shim
    from boardlaw_b200.mcts import cuda as mcts_shim
    ref = refpy.REFERENCE / 'boardlaw'
    for rel, planted in (('hex/cuda.py', hex_shim.hexcuda), ('mcts/cuda.py', mcts_shim.mctscuda)):
        used = set(re.findall(r'module\(\)\.(\w+)', (ref / rel).read_text()))
        assert used, f'no native call found in {rel}'
        missing = [name for name in used if not hasattr(planted, name)]
        assert not missing, f'the object planted for boardlaw/{rel} lacks {missing}'
    # the argument lists the reference's wrappers pass (boardlaw/mcts/cuda.py:26-42, boardlaw/hex/cuda.py:13-17)
    assert list(inspect.signature(mcts_shim.mctscuda.MCTS).parameters)[:7] == ['logits', 'w', 'n', 'c_puct', 'seats', 'terminal', 'children']
    assert {'v', 'w', 'n', 'rewards', 'parents', 'terminal'} <= set(inspect.signature(mcts_shim.mctscuda.Backup).parameters)
    assert len(inspect.signature(hex_shim.hexcuda.step).parameters) == 3 and len(inspect.signature(hex_shim.hexcuda.observe).parameters) == 2


@pytest.mark.gpu
@pytest.mark.reference
@needs_ref
@pytest.mark.parametrize('S,B,T', [(5, 64, 16), (9, 48, 32)])
def test_reference_python_on_these_kernels(S, B, T):
    """The reference's own Hex + MCTS, unmodified, on cuda through install(): same tree as on its own CPU kernels, given the
    same network outputs (evaluated once on the CPU, in fp32) and the same random numbers."""
    import boardlaw_b200.cuda
    from oracle import pyref
    r = refpy.load('O2')
    cpu_caches = (r.hex.cuda._cache, r.mcts.cuda._cache)
    sd = pyref.synth_state_dict(S, 32, 2, seed=21)
    onet = pyref.FCNet(sd)

    def network(world):                                   # network(worlds) -> arrdict(logits, v), on the world's device
        w = pyref.HexWorld(world.board.cpu(), world.seats.cpu().int())
        out = onet(w)
        return r.arrdict.arrdict(logits=out.logits.to(world.board.device), v=out.v.to(world.board.device))

    def search(device):
        torch.manual_seed(5)
        worlds = r.Hex.initial(B, S, device=device)
        g = torch.Generator().manual_seed(6)
        for _ in range(S * S // 3):
            a = torch.multinomial(worlds.valid.float().cpu(), 1, generator=g).squeeze(-1).to(device)
            worlds, _ = worlds.step(a)
        m = r.MCTS(worlds, n_nodes=T)
        # the reference draws its Dirichlet noise and descent random numbers on the tree's device: feed both runs the CPU streams
        m.initialize(network)
        if device != 'cpu':
            return worlds, m
        return worlds, m

    # CPU run on the reference's own kernels
    w_cpu, m_cpu = search('cpu')
    noise_row = m_cpu.decisions.logits[:, 0].clone()
    try:
        boardlaw_b200.cuda.install()
        w_gpu, m_gpu = search('cuda')
        assert torch.equal(w_gpu.board.cpu(), w_cpu.board) and torch.equal(w_gpu.seats.cpu(), w_cpu.seats)
        m_gpu.decisions.logits[:, 0] = noise_row.cuda()   # same noised root prior on both
        for sim in range(1, T):
            state = torch.get_rng_state()
            # both descents must see the same random numbers: the CUDA shim takes them injected, the CPU kernels draw them
            rands = torch.rand((B, T), dtype=torch.half)
            torch.set_rng_state(state)
            r.hex.cuda._cache, r.mcts.cuda._cache = cpu_caches
            m_cpu.simulate(network)
            boardlaw_b200.cuda.install()
            from boardlaw_b200.mcts import cuda as shim
            shim.inject_rands(rands.cuda())
            m_gpu.simulate(network)
            for name in ('children', 'parents', 'relation'):
                assert torch.equal(getattr(m_gpu.tree, name).cpu(), getattr(m_cpu.tree, name)), f'{name} differs at sim {sim}'
            assert torch.equal(m_gpu.stats.n.cpu(), m_cpu.stats.n)
            assert torch.equal(m_gpu.stats.w.cpu().view(torch.int16), m_cpu.stats.w.view(torch.int16))
        assert torch.equal(m_gpu.root().logits.cpu().view(torch.int16), m_cpu.root().logits.view(torch.int16))
    finally:
        r.hex.cuda._cache, r.mcts.cuda._cache = cpu_caches

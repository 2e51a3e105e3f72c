"""CPU-side tests (no GPU): the C-ABI library loads and exports every declared symbol, host logic of the fusion
plan (subset enumeration, bit-exact selection ranges), state_dict layout, weight-packing index maps, and the
data-parallel host path on a 2-rank gloo group."""
import os
import re
import subprocess
import sys

import pytest
import torch

from oracle import mopoe_oracle as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_loads_and_exports_every_declared_symbol():
    from mopoe_mimic_b200 import _lib as L
    from mopoe_mimic_b200 import build as B
    B.build()
    lib = L.load()                                   # binds every name in SIGNATURES (AttributeError otherwise)
    header = open(os.path.join(ROOT, 'include', 'mopoe_b200.h')).read()
    declared = set(re.findall(r'\b(mopoe_[a-z0-9_]+)\s*\(', header)) - {'mopoe_window_t', 'mopoe_rows_t'}
    assert declared, 'no declarations found'
    for name in sorted(declared):
        assert hasattr(lib, name), 'libmopoe_b200.so does not export %s' % name
    assert declared <= set(L.SIGNATURES), sorted(declared - set(L.SIGNATURES))
    assert lib.mopoe_version() >= 100


def test_product_refuses_cpu():
    import mopoe_mimic_b200 as P
    from mopoe_mimic_b200.engine import Engine
    with pytest.raises(RuntimeError):
        Engine('cpu')
    fl = P.default_flags(device=torch.device('cpu'), DIM_img=8, DIM_text=8, class_dim=16, batch_size=2)
    exp = P.Experiment(fl)
    with pytest.raises(RuntimeError):
        exp.mm_vae({'PA': torch.rand(2, 1, 128, 128), 'Lateral': torch.rand(2, 1, 128, 128),
                    'text': torch.zeros(2, 1024, 71)})


def test_product_does_not_import_oracle():
    src = os.path.join(ROOT, 'mopoe_mimic_b200')
    for fn in os.listdir(src):
        if fn.endswith('.py'):
            text = open(os.path.join(src, fn)).read()
            assert not re.search(r'^\s*(from|import)\s+[\w.]*oracle', text, re.M), fn


@pytest.mark.parametrize('kw', [dict(), dict(mods=('PA', 'text')), dict(img_size=64), dict(img_size=256)])
def test_state_dict_layout_matches_oracle_spec(kw):
    import mopoe_mimic_b200 as P
    small = dict(DIM_img=8, DIM_text=8, class_dim=16, batch_size=2)
    fl = P.default_flags(device=torch.device('cpu'), **small, **kw)
    sd = P.Experiment(fl).mm_vae.state_dict()
    spec = O.param_spec(O.default_flags(**small, **kw))
    assert [(k, tuple(v.shape)) for k, v in sd.items()] == [(k, tuple(s)) for k, s in spec.items()]


def test_set_subsets_and_selection_bit_exact():
    from mopoe_mimic_b200.fusion import FusionPlan, selection_ends, set_subsets, uniform_weights
    mods = {'PA': object(), 'Lateral': object(), 'text': object()}
    assert list(set_subsets(mods).keys()) == list(O.subset_keys(mods.keys()).keys())
    assert list(set_subsets({'PA': 1, 'text': 2}).keys()) == ['', 'PA', 'text', 'PA_text']
    for B in (1, 5, 6, 7, 8, 16, 37, 64, 128, 255, 256, 1000, 1024, 2048, 4096):
        for S in (1, 2, 3, 4, 7):
            assert selection_ends(B, uniform_weights(S)) == O.selection_bounds(B, [1.0 / S] * S)[1], (B, S)
    sub = O.subset_keys(('PA', 'Lateral', 'text'))
    for method, S in (('joint_elbo', 7), ('moe', 3), ('poe', 1)):
        plan = FusionPlan(['PA', 'Lateral', 'text'], ['PA', 'Lateral', 'text'], list(sub.keys()), list(sub.values()),
                          method, 256, 128, 256)
        assert plan.keys == [k for k in sub if k] and len(plan.stacked) == S
        # members are kept in the reference's stacking order (sorted by name): Lateral < PA < text
        assert [plan.cfg.mem_idx[3][j] for j in range(2)] == [1, 0]          # 'Lateral_PA' -> (Lateral, PA)
    # unimodal pass of the poe loss: only subsets fully present are built
    plan = FusionPlan(['PA', 'Lateral', 'text'], ['text'], list(sub.keys()), list(sub.values()), 'poe', 8, 32, 8)
    assert plan.keys == ['text'] and plan.cfg.prior_expert == 1


def test_weight_packing_index_maps_against_conv_identities():
    """conv-form / phase-form / full-form (torch reference re-layouts in engine.py, mirrored by the CUDA pack kernel):
    a stride-2 deconv equals its 4 sub-pixel phase GEMMs, checked on CPU with plain matmuls."""
    import torch.nn.functional as F
    from mopoe_mimic_b200.engine import KTAPS, conv_form, full_form, phase_form
    g = torch.Generator().manual_seed(0)
    ci, co, H = 3, 5, 4
    x = torch.randn(2, ci, H, H, generator=g)
    wt = torch.randn(ci, co, 4, 4, generator=g)
    ref = F.conv_transpose2d(x, wt, stride=2, padding=1)
    xp = F.pad(x, (1, 1, 1, 1)).permute(0, 2, 3, 1)                  # bordered channels-last
    ph = phase_form(wt, torch.float32)
    out = torch.zeros(2, 2 * H, 2 * H, co)
    for py in range(2):
        for px in range(2):
            for t in range(H):
                for s in range(H):
                    win = xp[:, t + py:t + py + 2, s + px:s + px + 2, :].reshape(2, -1)    # (r, kxi, ci)
                    out[:, 2 * t + py, 2 * s + px] = win @ ph[py * 2 + px].t()
    torch.testing.assert_close(out.permute(0, 3, 1, 2), ref, rtol=1e-5, atol=1e-5)
    w = torch.randn(co, ci, 4, 4, generator=g)
    refc = F.conv2d(x, w, stride=2, padding=1)
    wc = conv_form(w, torch.float32)
    outc = torch.zeros(2, H // 2, H // 2, co)
    for oy in range(H // 2):
        for ox in range(H // 2):
            win = xp[:, 2 * oy:2 * oy + 4, 2 * ox:2 * ox + 4, :].reshape(2, -1)
            outc[:, oy, ox] = win @ wc.t()
    torch.testing.assert_close(outc.permute(0, 3, 1, 2), refc, rtol=1e-5, atol=1e-5)
    z = torch.randn(2, ci, 1, 1, generator=g)
    reff = F.conv_transpose2d(z, wt, stride=1, padding=0)
    outf = (z.reshape(2, ci) @ full_form(wt, torch.float32).t()).reshape(2, 4, 4, co)
    torch.testing.assert_close(outf.permute(0, 3, 1, 2), reff, rtol=1e-5, atol=1e-5)
    assert KTAPS == ((3, 1), (2, 0))


_DP_WORKER = r'''
import os, sys, torch, torch.distributed as dist
sys.path.insert(0, %r)
from mopoe_mimic_b200.dp import FlatGradAllReduce, broadcast_flat, shard_batch, bucket_bounds
rank, world = int(os.environ['RANK']), int(os.environ['WORLD_SIZE'])
dist.init_process_group('gloo', init_method='tcp://127.0.0.1:%%s' %% os.environ['MASTER_PORT'], rank=rank, world_size=world)
torch.manual_seed(0)
full = {'PA': torch.arange(8 * 3, dtype=torch.float32).reshape(8, 3), 'text': torch.arange(8, dtype=torch.float32).reshape(8, 1)}
mine = shard_batch(full, rank, world)
assert mine['PA'].shape[0] == 8 // world and float(mine['text'][0]) == rank * (8 // world)
# "gradient" of a mean-over-local-batch loss: DP mean of per-rank grads == gradient of the global-batch mean
w = torch.ones(3, requires_grad=True)
loss = (mine['PA'] @ w).mean()
loss.backward()
n = 1000
flat = torch.zeros(n)
flat[:3] = w.grad
flat[3:] = rank + 1.0
ar = FlatGradAllReduce(bucket_mb=0.001)            # forces several buckets
assert len(bucket_bounds(n, ar.bucket_elems)) > 1
ar(flat)
flat *= ar.grad_scale
wg = torch.ones(3, requires_grad=True)
(full['PA'] @ wg).mean().backward()
assert torch.allclose(flat[:3], wg.grad), (flat[:3], wg.grad)
assert torch.allclose(flat[3:], torch.full((n - 3,), (1.0 + world) / 2))
p = torch.full((10,), float(rank))
broadcast_flat(p, 0)
assert float(p.sum()) == 0.0
# what train_step / GraphedTrainStep do when handed an all-reduce (the reference wraps the model in DDP): the optimizer
# takes the 1/world gradient scale and every rank starts from rank 0's parameters
from types import SimpleNamespace
from mopoe_mimic_b200.train import attach_allreduce
opt = SimpleNamespace(grad_scale=1.0)
exp = SimpleNamespace(optimizer=opt, mm_vae=SimpleNamespace(flat_params=torch.full((6,), float(rank + 1))))
ar2 = FlatGradAllReduce()
attach_allreduce(exp, ar2)
assert opt.grad_scale == 1.0 / world and float(exp.mm_vae.flat_params.sum()) == 6.0
exp.mm_vae.flat_params.add_(rank)            # idempotent: a second attach must not broadcast again
attach_allreduce(exp, ar2)
assert float(exp.mm_vae.flat_params[0]) == 1.0 + rank
dist.destroy_process_group()
print('rank', rank, 'ok')
'''


def test_data_parallel_host_path_gloo_world2(tmp_path):
    script = tmp_path / 'dp_worker.py'
    script.write_text(_DP_WORKER % ROOT)
    port = str(29500 + os.getpid() % 500)
    procs = []
    for r in range(2):
        env = dict(os.environ, RANK=str(r), WORLD_SIZE='2', MASTER_ADDR='127.0.0.1', MASTER_PORT=port)
        procs.append(subprocess.Popen([sys.executable, str(script)], env=env, stdout=subprocess.PIPE,
                                      stderr=subprocess.STDOUT, text=True))
    outs = [p.communicate(timeout=120)[0] for p in procs]
    for r, (p, o) in enumerate(zip(procs, outs)):
        assert p.returncode == 0, o
        assert 'rank %d ok' % r in o


def test_bench_reference_arm_prints_the_contract_line():
    """`bench.py --impl reference` (the CPU arm the driver runs first): one JSON line with the contract's keys, runnable
    without a GPU.  Tiny sample here (batch 2, one timed step)."""
    import json
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    have_ref = os.path.isdir(os.path.join(root, 'baseline', '_ref', 'mimic', 'networks'))
    for kind in (['reference'] if have_ref else []) + ['port']:
        r = subprocess.run([sys.executable, os.path.join(root, 'bench.py'), '--impl', 'reference', '--steps', '1', '--warmup', '1',
                            '--cpu-batch', '2', '--cpu-kind', kind], capture_output=True, text=True, timeout=600, cwd=root)
        assert r.returncode == 0, r.stderr[-2000:]
        line = json.loads(r.stdout.strip().splitlines()[-1])
        assert line['cpu_baseline']['kind'] == kind
    for k in ('metric', 'value', 'unit', 'n_gpus', 'steps', 'warmup', 'ms_per_step', 'higher_is_better', 'scaling',
              'vs_baseline', 'dtype', 'data', 'config', 'impl', 'cpu_baseline', 'e2e'):
        assert k in line, k
    assert line['impl'] == 'reference' and line['unit'] == 'samples/s' and line['value'] > 0
    assert line['cpu_baseline']['cores'] >= 1
    assert line['e2e']['h2d_bytes_per_step'] == 0 and line['e2e']['d2h_bytes_per_step'] == 0
    assert 'workload' in line['config']


def test_ctypes_structs_match_the_c_header(tmp_path):
    """The ctypes mirrors in _lib.py must have the same size and field offsets as the structs of include/mopoe_b200.h
    (compiled here with gcc): a silent mismatch would corrupt every call through the ABI."""
    import ctypes as C
    from mopoe_mimic_b200 import _lib as L
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    pairs = [('mopoe_view_t', L.View, ['ptr', 'dtype', 'B', 'C', 'ph', 'sB', 'sW']),
             ('mopoe_window_t', L.Window, ['a', 'E0', 'R', 'KW', 'a_off', 'sAr']),
             ('mopoe_rows_t', L.Rows, ['d', 'N', 'd_off', 's2']),
             ('mopoe_fusion_cfg_t', L.FusionCfg, ['M', 'fuse_mode', 'members', 'stacked', 'sel_end', 'mem_cnt', 'mem_idx',
                                                  'mem_end', 'norm']),
             ('mopoe_pack_job_t', L.PackJob, ['W', 'dst', 'A', 'bpad', 'tile0', 'nx']),
             ('mopoe_dp_peers_t', L.DpPeers, ['grad', 'param', 'flags']),
             ('mopoe_bn_req_t', L.BnReq, ['out', 'mask', 'mask_mode', 'nchunk', 'ws', 'ws_doubles', 'eps', 'momentum', 'mean',
                                          'running_var']),
             ('mopoe_res_req_t', L.ResReq, ['r', 'mean', 'invstd', 'gamma', 'beta', 'a', 'b', 'mask', 'mask_mode', 'out']),
             ('mopoe_bnbwd_req_t', L.BnBwdReq, ['x', 'mask', 'mask_mode', 'accumulate', 'mean', 'invstd', 'gamma', 'beta', 'ws',
                                                'ws_doubles', 'dgamma', 'dbeta', 'sums'])]
    src = ['#include <stdio.h>', '#include <stddef.h>', '#include "mopoe_b200.h"', 'int main(void) {']
    for cname, _, fields in pairs:
        src.append('  printf("%s %%zu", sizeof(%s));' % (cname, cname))
        for f in fields:
            src.append('  printf(" %%zu", offsetof(%s, %s));' % (cname, f))
        src.append('  printf("\\n");')
    src += ['  return 0;', '}']
    cfile = tmp_path / 'abi.c'
    cfile.write_text('\n'.join(src))
    exe = tmp_path / 'abi'
    subprocess.run(['gcc', '-I', os.path.join(root, 'include'), str(cfile), '-o', str(exe)], check=True)
    out = subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout.strip().splitlines()
    for line, (cname, ct, fields) in zip(out, pairs):
        parts = line.split()
        assert parts[0] == cname
        assert int(parts[1]) == C.sizeof(ct), (cname, parts[1], C.sizeof(ct))
        for f, off in zip(fields, parts[2:]):
            assert int(off) == getattr(ct, f).offset, (cname, f, off, getattr(ct, f).offset)


def test_jsd_plan_has_a_prior_component_and_exact_ranges():
    """jsd mode: the joint mixture has M+1 components (unimodal posteriors + the N(0,I) prior, BaseMMVae.py:180-186); the
    host plan marks the prior with -1 and its batch-row ranges equal the oracle's selection bounds bit for bit."""
    from mopoe_mimic_b200.fusion import FusionPlan
    from oracle import mopoe_oracle as O
    for mods, B in ((('PA', 'Lateral', 'text'), 256), (('PA', 'text'), 9), (('PA', 'Lateral', 'text'), 17)):
        keys = list(O.subset_keys(mods).keys())
        members = [sorted(k.split('_')) if k else [] for k in keys]
        plan = FusionPlan(list(mods), list(mods), keys, members, 'jsd', B, 32, B)
        uni = [i for i, k in enumerate(plan.keys) if '_' not in k]
        assert plan.stacked == uni + [-1]
        assert plan.cfg.S == len(mods) + 1 and plan.cfg.fuse_mode == 1 and plan.cfg.prior_expert == 0
        _, ends = O.selection_bounds(B, [1.0 / (len(mods) + 1)] * (len(mods) + 1))
        assert [plan.cfg.sel_end[j] for j in range(plan.cfg.S)] == ends == plan.sel_end
        assert [plan.cfg.stacked[j] for j in range(plan.cfg.S)] == plan.stacked


def test_peer_exchange_slices_cover_the_buffer():
    """owner slices of the fused exchange kernel: contiguous, float4-aligned, covering every element exactly once
    (the same formula as dp_adam_exchange_kernel: per = ceil(n/4 / world) float4s)"""
    from mopoe_mimic_b200.dp import PeerExchange

    class _Stub:
        slice_bounds = PeerExchange.slice_bounds

        def __init__(self, n, world):
            self.params = torch.empty(n)
            self.world = world
    for n, world in ((153067136, 8), (3145920, 2), (1024, 16), (64, 3), (4, 8)):
        b = _Stub(n, world).slice_bounds()
        assert len(b) == world and b[0][0] == 0 and b[-1][1] == n
        for (s0, e0), (s1, e1) in zip(b, b[1:]):
            assert e0 == s1 and s0 % 4 == 0 and e0 % 4 == 0 and s0 <= e0


def test_word_encoding_state_dict_matches_the_reference_layout():
    """word-encoded text (Embedding + Conv1d stem + 8 constructed / 6 used blocks; one nn.Sequential generator with a
    pointwise vocabulary head at len_sequence 128): names, order and shapes equal the oracle's spec, which is pinned
    against the reference's own state_dict (oracle/gen_golden.py asserts key order on the live model)."""
    import mopoe_mimic_b200 as P
    from oracle import mopoe_oracle as O
    for L_ in (128, 1024):
        kw = dict(batch_size=4, DIM_img=8, DIM_text=16, class_dim=16, text_encoding='word', vocab_size=48, len_sequence=L_)
        spec = O.param_spec(O.default_flags(**kw))
        exp = P.Experiment(P.default_flags(device=torch.device('cpu'), **kw))
        sd = exp.mm_vae.state_dict()
        assert [(k, tuple(v.shape)) for k, v in sd.items()] == [(k, tuple(s_)) for k, s_ in spec.items()]


@pytest.mark.parametrize('nd,pad', [(2, 0), (2, 1), (1, 0), (1, 1)])
def test_phase_rows_partition_the_interior_of_a_deconv_output(nd, pad):
    """Engine.phase_rows: the 2^nd sub-pixel phase row addressings of a stride-2 deconv output (plain or bordered) visit every
    interior pixel exactly once and nothing else — the addressing the phase-batched GEMM, its residual stream (the shortcut
    tensor) and its BatchNorm-backward stream (the BatchNorm input) all share."""
    import torch
    from mopoe_mimic_b200.engine import Act, Engine
    B, H, W, n = 3, (1 if nd == 1 else 4), 5, 8
    OH, OW = (1 if nd == 1 else 2 * H), 2 * W
    ph, pw = (0, pad) if nd == 1 else (pad, pad)
    act = Act(torch.zeros(B, OH + 2 * ph, OW + 2 * pw, n), B, OH, OW, n, ph, pw)
    seen = []
    for py in range(1 if nd == 1 else 2):
        for px in range(2):
            r = Engine.phase_rows(act, n, py, px)
            assert r.N == n and r.d == act.t.data_ptr()
            for b in range(B):
                for y in range(H):
                    for x in range(W):
                        seen.append(r.d_off + x * r.s0 + y * r.s1 + b * r.s2)
    want = [((b * act.Hs + ph + y) * act.Ws + pw + x) * n for b in range(B) for y in range(OH) for x in range(OW)]
    assert sorted(seen) == sorted(want) and len(set(seen)) == len(seen)

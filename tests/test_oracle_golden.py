"""The CPU oracle against the committed golden fixtures (generated from the live reference by
oracle/gen_golden.py).  Tolerance: fp64 oracle vs fp64 reference, 1e-9 relative (observed 1e-13)."""
import glob
import os

import pytest
import torch

from oracle import mopoe_oracle as O

CASES = sorted(os.path.basename(p)[:-3] for p in glob.glob(os.path.join(os.path.dirname(__file__), 'golden', '*.pt')))
FAST = [c for c in CASES if c.startswith('small') and '256' not in c]


def load(golden_dir, name):
    return torch.load(os.path.join(golden_dir, name + '.pt'), weights_only=False)


def run_oracle(fx, dtype=torch.float64):
    fl = O.default_flags(**fx['flags'])
    if 'rec_weights' not in fx['flags']:
        fl.rec_weights = {m: 0.33 for m in fl.mods}
    B = fx['actual_batch']
    st = O.make_state(fl, fx['seeds']['state'], dtype)
    batch = O.make_batch(fl, fx['seeds']['batch'], dtype, B)
    noise = [O.make_noise(fl, fx['seeds']['noise'] + i, dtype, B) for i in range(1 + len(fl.mods))]
    uni = {m: noise[1 + i] for i, m in enumerate(fl.mods)}
    es = O.make_style_noise(fl, fx['seeds']['noise'], dtype, B)
    ues = {m: O.make_style_noise(fl, fx['seeds']['noise'] + 1 + i, dtype, B) for i, m in enumerate(fl.mods)} if es is not None else None
    return fl, st, O.step_with_grads(st, batch, fl, noise[0][0], noise[0][1], uni_masks=uni, eps_style=es, uni_eps_style=ues)


def check_checksum(t, cs, rtol, scale=0.0):
    t = t.detach().double().reshape(-1)
    assert t.numel() == cs['numel']
    ref_l2 = cs['l2']
    assert abs(float(t.norm()) - ref_l2) <= rtol * (ref_l2 + scale)
    tol = rtol * (float(cs['val'].abs().max()) + ref_l2 / max(cs['numel'], 1) ** 0.5 + scale)
    assert float((t[cs['pos']] - cs['val']).abs().max()) <= tol


def test_state_spec_matches_reference(golden_dir):
    fx = load(golden_dir, 'cfg1_tri_128_b16_joint')
    spec = O.param_spec(O.default_flags())
    assert [(k, tuple(v)) for k, v in spec.items()] == [(k, tuple(s)) for k, s in fx['state_keys']]
    n_param = sum(int(torch.tensor(s).prod()) if len(s) else 1 for k, s in fx['state_keys']
                  if 'running_' not in k and 'num_batches' not in k)
    assert n_param == 153066953          # SURVEY.md §0 [probe]
    assert len(fx['state_keys']) == 744


def test_subset_order_bit_exact(golden_dir):
    assert list(O.subset_keys(('PA', 'Lateral', 'text'))) == \
        ['', 'PA', 'Lateral', 'text', 'Lateral_PA', 'PA_text', 'Lateral_text', 'Lateral_PA_text']
    assert list(O.subset_keys(('PA', 'text'))) == ['', 'PA', 'text', 'PA_text']
    for name in ('small_tri_joint', 'small_patext_joint'):
        fx = load(golden_dir, name)
        fl = O.default_flags(**fx['flags'])
        assert list(O.subset_keys(fl.mods)) == fx['subset_keys']


def test_selection_bounds_probe_values():
    # SURVEY.md §8 a11 [probe]
    assert O.selection_bounds(16, [1 / 7] * 7)[1] == [2, 4, 6, 8, 10, 12, 16]
    assert O.selection_bounds(256, [1 / 7] * 7)[1] == [36, 72, 108, 144, 180, 216, 256]
    assert O.selection_bounds(256, [1 / 3] * 3)[1] == [85, 170, 256]


@pytest.mark.parametrize('name', FAST)
def test_oracle_matches_reference_golden(golden_dir, name):
    fx = load(golden_dir, name)
    fl, st, out = run_oracle(fx)
    rt = 1e-9
    assert abs(float(out['total_loss']) - fx['total_loss']) <= rt * abs(fx['total_loss'])
    assert abs(float(out['results']['joint_divergence']) - fx['joint_divergence']) <= rt * abs(fx['joint_divergence'])
    for k, v in fx['klds'].items():
        assert abs(float(out['klds'][k]) - v) <= rt * abs(v)
    for k, v in fx['log_probs'].items():
        assert abs(float(out['log_probs'][k]) - v) <= rt * abs(v)
    lat = out['results']['latents']
    assert list(lat['subsets'].keys()) == list(fx['subsets'].keys())
    for k, (mu, lv) in fx['subsets'].items():
        torch.testing.assert_close(lat['subsets'][k][0], mu, rtol=rt, atol=1e-12)
        torch.testing.assert_close(lat['subsets'][k][1], lv, rtol=rt, atol=1e-12)
    torch.testing.assert_close(lat['mus'], fx['mus'], rtol=rt, atol=1e-12)
    torch.testing.assert_close(lat['joint'][0], fx['joint'][0], rtol=rt, atol=1e-12)
    for m, cs in fx['rec'].items():
        check_checksum(out['results']['rec'][m], cs, rt)
    assert set(out['grads'].keys()) == set(fx['grads'].keys())
    for k, cs in fx['grads'].items():
        check_checksum(out['grads'][k], cs, 1e-8, scale=1e-6 * fx['grad_scale'])
    for k, cs in fx['bn'].items():
        if k in out['results']['bn_updates']:      # (BN layers the step never runs keep their initial statistics)
            check_checksum(out['results']['bn_updates'][k], cs, rt)
        else:
            check_checksum(st[k], cs, rt)


def test_oracle_fp32_close_to_fp64_reference(golden_dir):
    """How far the reference's own fp32 arithmetic sits from fp64: sets the floor for fp32 GPU parity."""
    fx = load(golden_dir, 'small_tri_joint')
    fl, st, out = run_oracle(fx, torch.float32)
    assert abs(float(out['total_loss']) - fx['total_loss']) <= 1e-5 * abs(fx['total_loss'])
    for k, v in fx['klds'].items():
        assert abs(float(out['klds'][k]) - v) <= 1e-4 * abs(v)


def test_eval_path_oracle_matches_reference_fixture(golden_dir):
    """N2: eval-mode inference (full / partial batches), decode on B*K rows, per-sample log-probs, importance-sampled
    likelihoods and cond_generation of the oracle against the reference outputs in eval_tri.pt (fp64)."""
    from oracle import gen_golden_eval as GE
    fx = torch.load(os.path.join(golden_dir, 'eval_tri.pt'), weights_only=False)
    fl = O.default_flags(**fx['flags'])
    dt = torch.float64
    st = GE.eval_state(fl, dt)
    batch = O.make_batch(fl, seed=1, dtype=dt)
    noise = GE.eval_noise(fl, dt)
    K, B = fx['k_imp'], fl.batch_size
    with torch.no_grad():
        res = O.forward(st, batch, fl, None, None, train=False)
        for k, (mu, lv) in fx['subsets'].items():
            assert float((res['latents']['subsets'][k][0] - mu).abs().max()) < 1e-10
            assert float((res['latents']['subsets'][k][1] - lv).abs().max()) < 1e-10
        for tag, p in fx['partial'].items():
            o = O.forward(st, {m: batch[m][:p['rows']] for m in p['mods']}, fl, None, None, train=False, present=list(p['mods']))
            assert list(o['latents']['subsets'].keys()) == list(p['subsets'].keys())
            for k, (mu, lv) in p['subsets'].items():
                assert float((o['latents']['subsets'][k][0] - mu).abs().max()) < 1e-10, (tag, k)
            assert float((o['latents']['joint'][1] - p['joint'][1]).abs().max()) < 1e-10
        for s_key, d in fx['lhood'].items():
            mu, lv = fx['subsets'][s_key]
            z = (noise['eps_imp'] * torch.exp(0.5 * lv.unsqueeze(0)) + mu.unsqueeze(0)).view(K * B, -1)
            dec = O.decode(st, fl, z)
            ll = O.importance_likelihoods(fl, K, dec, batch, z, mu, lv)
            for k, v in d['ll'].items():
                assert abs(float(ll[k]) - v) <= 1e-10 * abs(v), (s_key, k)
            for m in fl.mods:
                cs = d['mean'][m]
                assert abs(float(O.likelihood_mean(m, dec[m]).double().norm()) - cs['l2']) <= 1e-10 * cs['l2']
        mu, lv = fx['subsets']['Lateral_PA_text']
        z = noise['eps_cg'] * torch.exp(0.5 * lv) + mu
        assert float((O.decode(st, fl, z)['PA'] - fx['cond_gen_pa_full']).abs().max()) < 1e-10

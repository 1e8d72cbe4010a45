"""CPU: the Light-CNN-29v2 host schedule (xfr_b200/lightcnn.py) driven through the torch emulation of the kernel set
reproduces the reference's outputs (tests/golden/lightcnn29v2_seed0.npz): firing order, (A, X) recipes, the Add closure
quirk, channel padding, the NHWC re-ordering of the fc weight, priors."""
import numpy as np
import pytest
import torch

from emul_backend import EmulBackend
from helpers import rel_err
from test_lightcnn_oracle import MODES, NUM_CLASSES, lc_setup
from xfr_b200 import synth
from xfr_b200.lightcnn import LightCNNEngine


def lc_inputs(G, imgs, noise):
    x = torch.cat((imgs[0:1], noise)).permute(0, 2, 3, 1).contiguous()             # [2,128,128,1]
    W2 = torch.from_numpy(np.concatenate((G['enc_mate'], G['enc_nonmate']))).float()
    return x, W2.unsqueeze(0).repeat(2, 1, 1).contiguous()


def unpad_split(P, c):
    """device layout [J,H,W,2*Cp] -> reference Split layout [J,2*c,H,W]"""
    cp = P.shape[-1] // 2
    return torch.cat((P[..., :c], P[..., cp:cp + c]), -1).permute(0, 3, 1, 2)


@pytest.mark.parametrize('mode,tag', MODES)
@pytest.mark.parametrize('impl', ['fp32', 'tf32x3'])
def test_all_modes(mode, tag, impl):
    G, sd, imgs, noise, _ = lc_setup()
    eng = LightCNNEngine(sd, EmulBackend(impl_name=impl))
    x, W2 = lc_inputs(G, imgs, noise)
    fc = eng.forward(imgs.permute(0, 2, 3, 1).contiguous())
    assert rel_err(fc[1:2].numpy(), G['enc_mate']) < 1e-5
    P1 = torch.zeros(2, 2)
    P1[:, 0] = 1
    m = eng.ebp(x, P1, W2, mode, saliency=False).clone().numpy()
    s = eng.ebp(x, P1, W2, mode).clone().numpy()
    c = eng.contrastive(x, W2, mode=mode).clone().numpy()
    t = eng.contrastive(x, W2, mode=mode, percentile=20).clone().numpy()
    assert m.shape == (2, 128, 128)
    # 'tf32x3' packs: the emulation multiplies relu(W) rounded to TF32 in the W+ GEMMs, as the product's two-pass plan does
    tol = 1e-5 if impl == 'fp32' else 1e-3
    for i, pname in enumerate(('smooth', 'noise')):
        assert rel_err(m[i], G['ebp_mwp_%s_%s' % (tag, pname)]) < tol
        assert rel_err(s[i], G['ebp_%s_%s' % (tag, pname)]) < tol
        assert rel_err(c[i], G['cebp_%s_%s' % (tag, pname)]) < 50 * tol
        assert rel_err(t[i], G['tcebp20_%s_%s' % (tag, pname)]) < 50 * tol


def test_recorded_P_and_hooked_fc2():
    G, sd, imgs, _, _ = lc_setup()
    eng = LightCNNEngine(sd, EmulBackend())
    x = imgs[0:1].permute(0, 2, 3, 1).contiguous()
    W2 = torch.from_numpy(np.concatenate((G['enc_mate'], G['enc_nonmate']))).float().unsqueeze(0)
    P1 = torch.zeros(1, 2)
    P1[0, 0] = 1
    eng.forward(x)
    P, names, P2 = eng.sweep().run(P1, W2, 'affineonly_with_prior', record=True)
    assert names == [str(k) for k in G['P_kinds']] and len(P) == 87
    sums = np.array([float(p.double().sum()) for p in P[:-1]])
    gs = G['Psum_awp_smooth'][:-1]
    assert np.max(np.abs(sums - gs) / (np.abs(gs) + 1e-30)) < 1e-5          # zero padding adds nothing to any firing
    assert unpad_split(P[-2], 48).shape == (1, 96, 128, 128)
    # the network's own fc2 as classifier: 88 firings
    Pk = torch.zeros(1, NUM_CLASSES)
    Pk[0, 5] = 1
    P, names, P2 = eng.sweep().run(Pk, sd['fc2.weight'], 'affineonly_with_prior', record=True, hooked_fc2=True)
    assert len(P) == 88 and names[0] == 'Linear'
    assert rel_err(P2.sum(-1)[0].numpy(), G['ebp_mwp_awp_fc2head']) < 1e-5
    c = eng.contrastive(x, sd['fc2.weight'], 5, 9, hooked_fc2=True, num_classes=NUM_CLASSES).clone().numpy()
    assert rel_err(c[0], G['cebp_awp_fc2head']) < 5e-4


def test_layerwise_prior_rows():
    """layerwise_ebp(mode='argmax') (whitebox.py:570-571) at several firings, batched as gradient rows with one prior each."""
    G, sd, imgs, _, _ = lc_setup()
    eng = LightCNNEngine(sd, EmulBackend())
    x = imgs[0:1].permute(0, 2, 3, 1).contiguous()
    W2 = torch.from_numpy(np.concatenate((G['enc_mate'], G['enc_nonmate']))).float().unsqueeze(0)
    P1 = torch.zeros(1, 2)
    P1[0, 0] = 1
    eng.forward(x)
    Pm, _, _ = eng.sweep().run(P1, W2, 'affineonly_with_prior', record=True)
    Pm = [None if p is None else p.clone() for p in Pm]
    ks = [int(k) for k in G['lw_k']]
    priors = {k: (r, (Pm[k] * (Pm[k] == Pm[k].max())).reshape(-1).contiguous()) for r, k in enumerate(ks)}
    _, _, P2 = eng.sweep().run(torch.zeros(len(ks), 2), W2, 'affineonly_with_prior', priors=priors)
    maps = P2.sum(-1).numpy()
    for r, k in enumerate(ks):
        assert rel_err(maps[r], G['lw_elem_%d' % k]) < 2e-5, k

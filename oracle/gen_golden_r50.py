"""Generate tests/golden/resnet50_128_real.npz by running the UNMODIFIED reference (TEST INFRASTRUCTURE).

Build container only (/root/reference present):   python oracle/gen_golden_r50.py

Real weights: the only real face-matcher weights bundled with the reference are the VGGFace2 ResNet-50-128d
(models/resnet50_128_pytorch/resnet50_128_pytorch.tar.gz).  They are extracted to oracle/_ref/resnet50_128.pth
(git-ignored, travels with the working tree to the GPU box; 95 MB) and never committed.  Real images: the bundled
VGGFace2 triplet data/n00000001_00000117.JPEG (probe) / n00000001_00000384.JPEG (mate) / n00000002_00000100.JPEG
(non-mate) and data/demo_face.jpg, pre-processed by the reference's own Whitebox_resnet50_128.preprocess
(whitebox.py:235-258); the 224x224 uint8 crops are stored in the golden file so the GPU box needs no JPEG.
"""
import os
import sys
import tarfile
import warnings

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
REF = '/root/reference'
sys.path[:0] = [os.path.join(HERE, 'shim'), REF + '/python', REF + '/models/resnet50_128_pytorch', ROOT]
warnings.filterwarnings('ignore')

import numpy as np  # noqa: E402
import PIL.Image  # noqa: E402
import torch  # noqa: E402

import resnet50_128  # noqa: E402  (the reference)
from xfr.models.whitebox import Whitebox, Whitebox_resnet50_128  # noqa: E402  (the reference)

MEAN = (131.0912, 103.8827, 91.4953)


def weights_path():
    dst = os.path.join(HERE, '_ref', 'resnet50_128.pth')
    if not os.path.exists(dst):
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        with tarfile.open(REF + '/models/resnet50_128_pytorch/resnet50_128_pytorch.tar.gz') as tf:
            member = [m for m in tf.getmembers() if m.name.endswith('resnet50_128.pth')][0]
            with tf.extractfile(member) as src, open(dst, 'wb') as out:
                out.write(src.read())
    return dst


def main():
    torch.set_num_threads(os.cpu_count())
    net = resnet50_128.resnet50_128(weights_path())
    net.eval()
    G = {}
    wbn = Whitebox_resnet50_128(net)
    files = dict(probe='n00000001_00000117.JPEG', mate='n00000001_00000384.JPEG', nonmate='n00000002_00000100.JPEG',
                 demo='demo_face.jpg')
    X = {}
    for k, f in files.items():
        x = wbn.preprocess(PIL.Image.open(os.path.join(REF, 'data', f)))
        crop = np.round(x[0].numpy().transpose(1, 2, 0).astype(np.float64) + np.array(MEAN)).astype(np.uint8)
        back = torch.from_numpy((crop.astype(np.float64) - np.array(MEAN)).transpose(2, 0, 1).astype(np.float32)).unsqueeze(0)
        assert torch.equal(back, x), k
        G['crop_' + k] = crop
        X[k] = x
    with torch.no_grad():
        x_mate = wbn.encode(X['mate']).detach()
        x_non = wbn.encode(X['nonmate']).detach()
        x_probe = wbn.encode(X['probe']).detach()
    G['enc_mate'], G['enc_nonmate'], G['enc_probe'] = x_mate.numpy(), x_non.numpy(), x_probe.numpy()
    cos = torch.nn.functional.cosine_similarity
    print('cos(probe,mate)=%.3f cos(probe,nonmate)=%.3f' % (float(cos(x_probe, x_mate)), float(cos(x_probe, x_non))))
    P0 = torch.zeros(1, 2)
    P0[0, 0] = 1.0
    for mode in ('affineonly_with_prior', 'all', 'norelu', 'affineonly'):
        tag = {'affineonly_with_prior': 'awp'}.get(mode, mode)
        torch.manual_seed(0)
        wb = Whitebox(Whitebox_resnet50_128(net), ebp_subtree_mode=mode)
        wb.net.set_triplet_classifier(x_mate, x_non)                     # un-scaled rows
        for pname in ('probe', 'demo'):
            G['ebp_mwp_%s_%s' % (tag, pname)] = wb.ebp(X[pname], P0, mwp=True)
            G['Psum_%s_%s' % (tag, pname)] = np.array([float(p.double().sum()) for p in wb.P])
            if mode == 'affineonly_with_prior' and pname == 'probe':
                G['P_kinds'] = np.array([n.split('(')[0] for n in wb.P_layername])
                G['P_numel'] = np.array([p.numel() for p in wb.P])
            G['ebp_%s_%s' % (tag, pname)] = wb.ebp(X[pname], P0)
            G['cebp_%s_%s' % (tag, pname)] = wb.contrastive_ebp(X[pname], 0, 1)
            G['tcebp20_%s_%s' % (tag, pname)] = wb.truncated_contrastive_ebp(X[pname], 0, 1, percentile=20)
        print(mode, 'cebp max %.6e argmax %d | tcebp max %.6e | ebp max %.6e argmax %d' % (
            G['cebp_%s_probe' % tag].max(), G['cebp_%s_probe' % tag].argmax(), G['tcebp20_%s_probe' % tag].max(),
            G['ebp_%s_probe' % tag].max(), G['ebp_%s_probe' % tag].argmax()))
    # layerwise_ebp (whitebox.py:561-581) and weighted_subtree_ebp (647-737) on the real triplet, default mode
    wb = Whitebox(Whitebox_resnet50_128(net))
    wb.net.set_triplet_classifier(x_mate, x_non)
    ks = (0, 3, 10, 45, 80, 120, 150, 155)
    G['lw_k'] = np.array(ks)
    wb.ebp(X['probe'], P0)
    G['lw_el_idx'] = np.array([int(torch.argmax(wb.P[k].flatten())) for k in ks])
    for k, e in zip(ks, G['lw_el_idx']):
        G['lw_argmax_%d' % k] = wb.layerwise_ebp(X['probe'], k_layer=k, mode='argmax', mwp=True)
        G['lw_el_%d' % k] = wb.layerwise_ebp(X['probe'], k_layer=k, mode='elementwise', k_element=int(e), mwp=True)
    wbs = Whitebox(Whitebox_resnet50_128(net))
    wbs.net.set_triplet_classifier(x_mate, x_non)
    smap, P_img, P_sub, k_sub = wbs.weighted_subtree_ebp(X['probe'], 0, 1, topk=16, verbose=False, do_max_subtree=False,
                                                         do_mated_similarity_gating=True, subtree_mode='affineonly_with_prior')
    G['ws_smap'], G['ws_scores'], G['ws_k'], G['ws_first'] = smap, np.array(P_sub, dtype=np.float64), np.array(k_sub), P_img[-1]
    print('weighted_subtree k', list(k_sub), 'scores', ['%.3g' % v for v in P_sub])
    out = os.path.join(ROOT, 'tests', 'golden', 'resnet50_128_real.npz')
    np.savez_compressed(out, **G)
    print('wrote', out, os.path.getsize(out) // 1024, 'KB')


if __name__ == '__main__':
    main()

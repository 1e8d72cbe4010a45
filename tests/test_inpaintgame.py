"""the inpainting-game front end and scoring (xfr_b200/inpaintgame.py) against what the
reference's own functions return (tests/golden/inpaintgame_seed0.npz: oracle/gen_golden_inpaintgame.py ran
python/xfr/inpainting_game/generate_whitebox_saliency.py:81-215 and inpainting_game.py:12-146 on the seeded inputs of
tests/inpaintgame_fixture.py), and against the reference's per-job flow restated on the batch-1 Whitebox API.
CPU: host logic over the kernel emulation; GPU: the CUDA kernels."""
import os

import numpy as np
import pytest
import torch

from helpers import L1111, rel_err
from inpaintgame_fixture import images as _images, jobs as _jobs
from test_layerwise_subtree import _net
from xfr_b200 import inpaintgame as IG
from xfr_b200 import synth, whitebox


def _reference_flow(wb, im_mates, im_nonmates, probe_im, truncate_percent):
    """generate_whitebox_saliency.py:92-116, line by line on the batch-1 API."""
    x_mates = [wb.encode(wb.convert_from_numpy(im)).detach() for im in im_mates]
    x_nonmates = [wb.encode(wb.convert_from_numpy(im)).detach() for im in im_nonmates]
    avg_x_mate = torch.mean(torch.stack(x_mates), axis=0)
    avg_x_mate /= torch.norm(avg_x_mate)
    avg_x_nonmate = torch.mean(torch.stack(x_nonmates), axis=0)
    avg_x_nonmate /= torch.norm(avg_x_nonmate)
    img_probe = wb.convert_from_numpy(probe_im)
    wb.net.set_triplet_classifier((1.0 / 2500.0) * avg_x_mate, (1.0 / 2500.0) * avg_x_nonmate)
    if truncate_percent is None:
        return wb.contrastive_ebp(img_probe, k_poschannel=0, k_negchannel=1)
    return wb.truncated_contrastive_ebp(img_probe, k_poschannel=0, k_negchannel=1, percentile=truncate_percent)


def _check_batch_matches_per_job(gpu, tol):
    wb = whitebox.Whitebox(_net(L1111, gpu))
    jobs = _jobs()
    for pct in (None, 20):
        want = [_reference_flow(wb, *j, truncate_percent=pct) for j in jobs]
        got = IG.run_contrastive_triplet_ebp_batch(wb, jobs, truncate_percent=pct)
        assert got.shape == (3, 112, 112) and got.dtype == np.float32
        for g, w in zip(got, want):
            assert rel_err(g, w) < tol and abs(float(g.sum()) - 1.0) < 1e-3
        one = IG.run_contrastive_triplet_ebp(wb, *jobs[1], net_name='resnetv4_pytorch', ebp_version=6, truncate_percent=pct)
        assert rel_err(one, want[1]) < tol
    assert rel_err(got[0], got[2]) > 0.1            # different triplets give different maps: rows were not mixed up


def _check_jobs_vs_reference(gpu, tol_c, tol):
    """The batched front end against the outputs of the reference's own run_contrastive_triplet_ebp /
    run_weighted_subtree_triplet_ebp / mean_ebp (batch 1, one fresh reference Whitebox per call)."""
    G = _gold()
    jobs = _jobs()
    wb = whitebox.Whitebox(_net(L1111, gpu))
    for pct, tag in ((None, ''), (20, '_pct20')):
        got = IG.run_contrastive_triplet_ebp_batch(wb, jobs, truncate_percent=pct)
        for i in range(3):
            ref = G['job%d_contrastive%s' % (i, tag)]
            assert got[i].shape == ref.shape == (112, 112) and float(np.abs(got[i] - ref).max()) < 1e-4     # the north-star bar
            assert rel_err(got[i], ref) < tol_c, (i, pct)
    im = _images(4, seed=21)
    wb = whitebox.Whitebox(_net(L1111, gpu))
    assert rel_err(IG.mean_ebp(wb, im[3]), G['mean_ebp']) < tol
    for ctor_mode, sub_mode, ver, key in (('norelu', 'all', 6, 'ws_eval_smap'),                       # the eval flow's settings
                                          ('affineonly_with_prior', 'affineonly_with_prior', 7, 'ws_v7_smap')):   # + max, gating
        wb = whitebox.Whitebox(_net(L1111, gpu), ebp_subtree_mode=ctor_mode)
        got = IG.run_weighted_subtree_triplet_ebp(wb, im[0:2], im[2:3], im[3], net_name='resnetv4_pytorch',
                                                  subtree_mode_weighted=sub_mode, ebp_version=ver, device=None, topk=4)
        assert got.shape == (112, 112) and np.isfinite(got).all() and abs(float(got.sum()) - 1.0) < 1e-3
        assert rel_err(got, G[key]) < tol


def test_jobs_vs_reference_emulated():
    _check_jobs_vs_reference(False, 1e-3, 2e-3)


@pytest.mark.gpu
def test_jobs_vs_reference_gpu():
    # tolerances of tests/test_gpu_parity.py: the contrastive maps of this synthetic net are cancellation-amplified (mate and
    # non-mate rows are encodings of random-weight images, cos = 0.9999: measured 7e-2 of the map maximum on the default bf16x2
    # plan, 3e-2 on the split-TF32 plan); the asserted parity bar is 1e-4 max-abs (measured <= 3e-6)
    _check_jobs_vs_reference(True, 1.5e-1, 1e-2)


def test_batch_matches_per_job_emulated():
    _check_batch_matches_per_job(False, 1e-4)


@pytest.mark.gpu
def test_batch_matches_per_job_gpu():
    # batch-1 and batch-3 sweeps tile the GEMM rows differently (4-D TMA boxes): agreement to split-TF32 rounding
    _check_batch_matches_per_job(True, 5e-3)


def _shard_worker(rank, world, port, q):
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    torch.set_num_threads(4)
    wb = whitebox.Whitebox(_net(L1111, False))
    jobs = _jobs()                                             # 3 jobs over 2 ranks: shards of 2 and 1
    got = IG.run_contrastive_triplet_ebp_sharded(wb, jobs, truncate_percent=None)
    # configs[3]: the weighted-subtree jobs sharded the same way (eval settings, small topk to keep the CPU test short)
    wbs = whitebox.Whitebox(_net(L1111, False), ebp_subtree_mode='norelu')
    ws = IG.run_weighted_subtree_triplet_ebp_sharded(wbs, jobs[:2], subtree_mode_weighted='all', topk=4)
    if rank == 0:
        want = IG.run_contrastive_triplet_ebp_batch(wb, jobs)
        ws_want = np.stack([IG.run_weighted_subtree_triplet_ebp(wbs, *j, subtree_mode_weighted='all', topk=4) for j in jobs[:2]])
        q.put((got.shape, float(np.abs(got - want).max() / want.max()), ws.shape, float(np.abs(ws - ws_want).max() / ws_want.max())))
    else:
        assert got is None and ws is None
    dist.destroy_process_group()


def test_sharded_jobs_world2_gloo():
    """SURVEY 8e for the front end: contiguous job shards, one gather of maps, job order preserved (ragged shards)."""
    import socket
    import torch.multiprocessing as mp
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    procs = [ctx.Process(target=_shard_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(600)
        assert p.exitcode == 0
    shape, err, ws_shape, ws_err = q.get(timeout=10)
    assert shape == (3, 112, 112) and err < 1e-4      # torch CPU GEMMs block a 2-probe and a 3-probe batch differently
    assert ws_shape == (2, 112, 112) and ws_err < 1e-6


def test_ragged_sweeps_emulated(monkeypatch):
    """Jobs that do not fill the last engine sweep: 3 jobs / 7 gallery images in sweeps of 2 equal one sweep of everything."""
    wb = whitebox.Whitebox(_net(L1111, False))
    jobs = _jobs()
    want = IG.run_contrastive_triplet_ebp_batch(wb, jobs, truncate_percent=20)
    monkeypatch.setattr(whitebox, '_CHUNK', 2)
    got = IG.run_contrastive_triplet_ebp_batch(wb, jobs, truncate_percent=20)
    assert got.shape == (3, 112, 112) and rel_err(got, want) < 1e-4
    with pytest.raises(ValueError):
        IG.run_contrastive_triplet_ebp_batch(wb, [([], jobs[0][1], jobs[0][2])])


def test_mean_encodings_and_errors():
    wb = whitebox.Whitebox(_net(L1111, False))
    im = _images(4, seed=5)
    rows = IG.mean_encodings(wb, [im[0:3], im[3:4]])
    assert rows.shape == (2, 512) and torch.allclose(torch.norm(rows, dim=1), torch.ones(2), atol=1e-6)
    single = wb.encode(wb.convert_from_numpy(im[3]))[0]
    assert torch.allclose(rows[1], single / torch.norm(single), atol=1e-6)
    with pytest.raises(ValueError):
        IG.mean_encodings(wb, [im[0:1], []])
    with pytest.raises(RuntimeError):
        IG.saliency_method_name(wb, 'rise', 6, 'cuda')


def test_method_names():
    """generate_whitebox_saliency.py:310-378; the evaluation globs these names (plot_inpainting_game.py)."""
    wb = whitebox.Whitebox(_net(L1111, False))
    assert IG.saliency_method_name(wb, 'meanEBP', 6, 'cuda') == 'meanEBP_mode=awp_v06_cuda'
    assert IG.saliency_method_name(wb, 'contrastive', 6, 'cuda') == 'contrastive_triplet_ebp_mode=awp_v06_cuda'
    assert IG.saliency_method_name(wb, 'contrastive', 6, 'cpu', truncate_percent=20) == \
        'trunc_contrastive_triplet_ebp_mode=awp_v06_pct20_cpu'
    wb2 = whitebox.Whitebox(_net(L1111, False), ebp_subtree_mode='norelu')
    assert IG.saliency_method_name(wb2, 'weighted-subtree', 6, 'cuda', topk=32, subtree_mode_weighted='all') == \
        'weighted_subtree_triplet_ebp_mode=norelu,all_v06_top32_cuda'


def test_process_saliency_and_npz(tmp_path):
    g = np.random.default_rng(0)
    smap = g.random((112, 112)).astype(np.float32) ** 4
    probe = np.zeros((224, 224, 3), np.uint8)
    out = IG.process_saliency(probe, smap)
    assert out.shape == (224, 224) and out.min() >= 0.0 and out.max() <= 1.0
    # cubic B-spline interpolation reproduces a linear ramp away from the zero-padded border
    ramp = np.tile(np.linspace(0, 1, 112, dtype=np.float64), (112, 1))
    r = IG.process_saliency(probe, ramp)
    centres = (np.arange(224) + 0.5) / 2 - 0.5                   # output pixel centres in input coordinates
    assert np.abs(r[100, 48:-48] - centres[48:-48] / 111).max() < 1e-6
    assert np.array_equal(IG.process_saliency(np.zeros((112, 112)), smap), (smap - smap.min()) / (smap.max() - smap.min() + 1e-9))
    f = os.path.join(str(tmp_path), 'sub', '00012-contrastive_triplet_ebp_mode=awp_v06_cuda-saliency.npz')
    stored = IG.save_smap(f, smap, probe)
    back = np.load(f)['saliency_map']                            # the key plot_inpainting_game.py:228 reads
    assert back.shape == (224, 224) and np.array_equal(back, stored)


def test_generate_wb_smaps_batch_emulated(tmp_path):
    wb = whitebox.Whitebox(_net(L1111, False))
    jobs = _jobs()[:2]
    dirs = [os.path.join(str(tmp_path), 'subject_ID_%d' % i) for i in range(2)]
    files = IG.generate_wb_smaps_batch(wb, jobs, dirs, ['00003', '00017'], ebp_ver=6, device=None)
    assert len(files) == 4 and all(os.path.exists(f) for f in files)
    assert os.path.basename(files[0]) == '00003-contrastive_triplet_ebp_mode=awp_v06_cpu-saliency.npz'
    assert os.path.basename(files[3]) == '00017-trunc_contrastive_triplet_ebp_mode=awp_v06_pct20_cpu-saliency.npz'
    m = np.load(files[1])['saliency_map']
    want = IG.save_smap(os.path.join(str(tmp_path), 'x.npz'), _reference_flow(wb, *jobs[1], truncate_percent=None), jobs[1][2])
    assert m.shape == (224, 224) and rel_err(m, want) < 1e-4
    assert IG.generate_wb_smaps_batch(wb, jobs, dirs, ['00003', '00017'], ebp_ver=6, overwrite=False) == []


def test_weighted_subtree_and_mean_ebp_wrappers_emulated():
    im = _images(4, seed=21)
    wb = whitebox.Whitebox(_net(L1111, False), ebp_subtree_mode='norelu')
    got = IG.run_weighted_subtree_triplet_ebp(wb, im[0:2], im[2:3], im[3], net_name='resnetv4_pytorch',
                                              subtree_mode_weighted='all', ebp_version=6, device=None, topk=4)
    # generate_whitebox_saliency.py:134-143, 192-199 on the batch-1 API
    rows = IG.mean_encodings(wb, [im[0:2], im[2:3]])
    wb.net.set_triplet_classifier(rows[0:1], rows[1:2])
    want = wb.weighted_subtree_ebp(wb.convert_from_numpy(im[3]), 0, 1, topk=4, verbose=False, do_max_subtree=False,
                                   subtree_mode='all', do_mated_similarity_gating=False)[0]
    assert got.shape == (112, 112) and rel_err(got, want) < 1e-5
    assert wb.ebp_subtree_mode() == 'all'
    m = IG.mean_ebp(wb, im[3])
    assert m.shape == (112, 112) and abs(float(m.sum()) - 1.0) < 1e-3


# ------------------------------------------------------------------ row 3: scoring, pinned to the reference's outputs
# tests/golden/inpaintgame_seed0.npz: oracle/gen_golden_inpaintgame.py ran the reference's create_threshold_masks and
# classified_as_inpainted_twin (python/xfr/inpainting_game/inpainting_game.py) on tests/inpaintgame_fixture.py's inputs
def _gold():
    from helpers import GOLD
    return np.load(os.path.join(GOLD, 'inpaintgame_seed0.npz'))


def _digest(masks):
    import hashlib
    return hashlib.sha256(np.packbits(masks.astype(bool)).tobytes()).hexdigest()


def test_threshold_masks_bit_exact_vs_reference():
    from inpaintgame_fixture import PCT_DENSITY, PCT_PIXELS, scoring_fixture
    F, G = scoring_fixture(), _gold()
    cases = (('density', F['smap'], dict(threshold_method='percent-density', percentiles=PCT_DENSITY, seed=0)),
             ('density_nozero', F['smap_sparse'], dict(threshold_method='percent-density', percentiles=PCT_DENSITY, seed=3,
                                                       include_zero_elements=False)),
             ('pixels', F['smap'], dict(threshold_method='percent-pixels', percentiles=PCT_PIXELS, seed=1)),
             ('thresholds', F['smap'], dict(threshold_method='thresholds', thresholds=np.array([0.5, 1e-4, 2e-5, 1e-5, 0.0]),
                                            percentiles=None, seed=2)))
    state = np.random.get_state()[1].copy()
    for tag, smap, kw in cases:
        m = IG.create_threshold_masks(smap, **kw)
        assert m.dtype == bool and m.shape[1:] == (224, 224)
        assert np.array_equal(m.reshape(m.shape[0], -1).sum(1), G['masks_%s_count' % tag]), tag
        assert _digest(m) == str(G['masks_%s_sha256' % tag]), tag
        value, thr = IG.mask_value_map(smap, **kw)                     # what the device kernel consumes
        assert value.dtype == thr.dtype == np.float64 and np.array_equal(value[None] > thr[:, None, None], m)
    assert np.array_equal(np.random.get_state()[1], state)             # the global generator is left alone
    assert not G['masks_density_count'][0] and G['masks_density_count'][-1] == 224 * 224       # 0 %: nothing, 100 %: everything
    mb = IG.create_threshold_masks(F['smap'], 'percent-density', percentiles=PCT_DENSITY[::10], seed=0, blur_sigma=4)
    assert mb.dtype == np.float32
    assert np.array_equal(mb[:, 100, :].astype(np.float64), G['masks_blur_row'])
    assert np.array_equal(mb.reshape(mb.shape[0], -1).astype(np.float64).sum(1), G['masks_blur_sum'])


def _check_scoring(gpu, tol):
    from inpaintgame_fixture import PCT_DENSITY, scoring_fixture
    F, G = scoring_fixture(), _gold()
    snet = whitebox.Whitebox(_net(L1111, gpu))
    gal_o, gal_p = snet.embeddings([F['orig']]), snet.embeddings([F['inp']])
    assert np.abs(gal_o - G['gal_orig']).max() < tol and np.abs(gal_p - G['gal_inp']).max() < tol
    cls, pg, pr = IG.classified_as_inpainted_twin(snet, F['orig'], F['inp'], G['gal_orig'], G['gal_inp'], F['smap'],
                                                  mask_threshold_method='percent-density', percentiles=PCT_DENSITY, seed=0)
    assert cls.shape == pg.shape == pr.shape == (101,) and cls.dtype == bool
    assert np.abs(pg - G['pg_dist']).max() < tol and np.abs(pr - G['pr_dist']).max() < tol
    margin = np.abs(G['pg_dist'] - G['pr_dist']) > 4 * tol               # blends the matcher does not place on the boundary
    assert np.array_equal(cls[margin], G['cls'][margin]) and margin.sum() >= 95
    # blurred masks (float32, explicit) and the plotting outputs
    pct = PCT_DENSITY[::10]
    cls, pg, pr, blends, masks = IG.classified_as_inpainted_twin(
        snet, torch.from_numpy(F['orig']), F['inp'], G['gal_orig'], G['gal_inp'], F['smap'], mask_threshold_method='percent-density',
        percentiles=pct, seed=0, mask_blur_sigma=4, return_transitions=True)
    assert np.abs(pg - G['pg_dist_blur']).max() < tol and np.abs(pr - G['pr_dist_blur']).max() < tol
    assert blends.shape == (11, 3, 224, 224) and blends.dtype == np.float64 and masks.shape == (11, 224, 224)
    assert np.array_equal(blends[0], F['orig'].astype(np.float64)) and np.array_equal(blends[-1], F['inp'].astype(np.float64))
    with pytest.raises(ValueError):
        IG.classified_as_inpainted_twin(snet, np.zeros((224, 224, 3)), np.zeros((224, 224, 3)), G['gal_orig'], G['gal_inp'],
                                        F['smap'], 'percent-density', percentiles=pct, seed=0)


def test_scoring_emulated():
    _check_scoring(False, 2e-6)


@pytest.mark.gpu
def test_scoring_gpu():
    _check_scoring(True, 2e-5)


@pytest.mark.gpu
def test_twin_blends_kernel_bit_exact():
    """xfrb_twin_blends against the numpy expression of inpainting_game.py:124-132 + .float(): thresholded and explicit
    masks (float64, and float32 whose 1 - m numpy rounds to fp32), 3-channel and 1-channel (Light-CNN) images."""
    from xfr_b200.kernels import CudaBackend
    dev = torch.device('cuda:0')
    be = CudaBackend(dev)
    rng = np.random.RandomState(0)
    for C, H in ((3, 224), (1, 128)):
        o = (rng.rand(C, H, H) * 255 - 100).astype(np.float32).astype(np.float64)
        p = (rng.rand(C, H, H) * 255 - 100)                                 # genuinely double-precision pixels
        value = rng.rand(H, H)
        thr = np.array([1.0, 0.9, 0.5, value[3, 5], 0.0])
        up = lambda a: torch.from_numpy(np.ascontiguousarray(a, dtype=np.float64)).to(dev)
        out = torch.empty(5, H, H, C, device=dev)
        be.twin_blends(up(o), up(p), up(value), up(thr), None, out)
        m = (value[None] > thr[:, None, None])[:, None]
        want = ((1.0 - m) * o[None] + m * p[None]).astype(np.float32)
        assert np.array_equal(out.cpu().numpy(), want.transpose(0, 2, 3, 1))
        for dt in (np.float64, np.float32):
            mk = rng.rand(4, H, H).astype(dt)
            mk[0], mk[1] = 0, 1
            be.twin_blends(up(o), up(p), None, None, up(mk), out[:4], mask_f32=(dt == np.float32))
            want = ((1.0 - mk[:, None]) * o[None] + mk[:, None] * p[None]).astype(np.float32)
            assert np.array_equal(out[:4].cpu().numpy(), want.transpose(0, 2, 3, 1)), (C, dt)


def test_scoring_lightcnn_one_channel_emulated():
    """The 1-channel path of the scoring (Light-CNN images are [1,128,128], inpainting_game.py:110-113): device blends + forward
    sweep against the numpy expression of inpainting_game.py:124-146 evaluated blend by blend on the same plugin."""
    from emul_backend import EmulBackend
    from xfr_b200.lightcnn import LightCNNEngine

    class _EmulLightCNN(whitebox.WhiteboxLightCNN):
        def _device(self):
            return torch.device('cpu')

        def engine(self, with_bias=False):
            if self._engine is None:
                self._engine = LightCNNEngine(self._sd, EmulBackend(), with_bias=with_bias)
            return self._engine

    snet = whitebox.Whitebox(_EmulLightCNN(synth.lightcnn_state_dict(0, 2)))
    rng = np.random.RandomState(5)
    orig, inp = rng.rand(1, 128, 128).astype(np.float32), rng.rand(1, 128, 128).astype(np.float32)
    smap = rng.rand(128, 128) ** 3
    pct = np.array([0, 10, 35, 70, 100])
    gal_o, gal_p = snet.embeddings([orig]), snet.embeddings([inp])
    cls, pg, pr = IG.classified_as_inpainted_twin(snet, orig, inp, gal_o, gal_p, smap, 'percent-density', percentiles=pct, seed=4)
    masks = IG.create_threshold_masks(smap, 'percent-density', percentiles=pct, seed=4)[:, np.newaxis]
    blends = (1.0 - masks) * orig.astype(np.float64)[np.newaxis] + masks * inp.astype(np.float64)[np.newaxis]
    emb = snet.embeddings(blends)
    emb = emb / np.linalg.norm(emb, axis=1, keepdims=True)
    assert np.abs(pr - np.linalg.norm(emb - gal_o, axis=1)).max() < 1e-6 and np.abs(pg - np.linalg.norm(emb - gal_p, axis=1)).max() < 1e-6
    assert not cls[0] and cls[-1] and pr[0] < 1e-6 and pg[-1] < 1e-6           # 0 %: the original probe; 100 %: the inpainted twin


def test_scoring_batch_matches_per_job_emulated(monkeypatch):
    """classified_as_inpainted_twin_batch == classified_as_inpainted_twin job by job (blends of several jobs in one sweep,
    sweeps cut across jobs)."""
    from inpaintgame_fixture import scoring_fixture
    F, G = scoring_fixture(), _gold()
    snet = whitebox.Whitebox(_net(L1111, False))
    pct = np.array([0, 20, 50, 80, 100])
    rng = np.random.RandomState(9)
    smap2 = (F['smap'] * (0.5 + rng.rand(224, 224))).astype(np.float32)
    jobs = [(F['orig'], F['inp'], G['gal_orig'], G['gal_inp'], F['smap']),
            (F['inp'], F['orig'], G['gal_inp'], G['gal_orig'], smap2),
            (torch.from_numpy(F['orig']), F['inp'], G['gal_orig'], G['gal_inp'], smap2)]
    want = [IG.classified_as_inpainted_twin(snet, j[0], j[1], j[2], j[3], j[4], 'percent-density', percentiles=pct, seed=s)
            for j, s in zip(jobs, (0, 1, 2))]
    monkeypatch.setattr(whitebox, '_CHUNK', 4)                    # 15 blends in sweeps of 4: sweeps straddle the jobs
    got = IG.classified_as_inpainted_twin_batch(snet, jobs, 'percent-density', percentiles=pct, seed=[0, 1, 2])
    assert len(got) == 3
    for (c0, g0, r0), (c1, g1, r1) in zip(want, got):
        assert np.array_equal(c0, c1) and np.abs(g0 - g1).max() < 1e-6 and np.abs(r0 - r1).max() < 1e-6
    assert IG.classified_as_inpainted_twin_batch(snet, [], 'percent-density', percentiles=pct) == []
    with pytest.raises(ValueError):
        IG.classified_as_inpainted_twin_batch(snet, jobs[:1] + [(np.zeros((1, 128, 128)),) * 2 + jobs[0][2:4] + (np.ones((128, 128)),)],
                                              'percent-density', percentiles=pct)


def test_front_end_lightcnn_emulated():
    """The batched front end through the Light-CNN plugin (grayscale 128x128 network input, 128x128 maps, 256-d encodings)."""
    from emul_backend import EmulBackend
    from xfr_b200.lightcnn import LightCNNEngine

    class _EmulLightCNN(whitebox.WhiteboxLightCNN):
        def _device(self):
            return torch.device('cpu')

        def engine(self, with_bias=False):
            if self._engine is None:
                self._engine = LightCNNEngine(self._sd, EmulBackend(), with_bias=with_bias)
            return self._engine

    wb = whitebox.Whitebox(_EmulLightCNN(synth.lightcnn_state_dict(0, 2)))
    jobs = _jobs()[:2]
    got = IG.run_contrastive_triplet_ebp_batch(wb, jobs)
    assert got.shape == (2, 128, 128) and got.dtype == np.float32
    for j, g in zip(jobs, got):
        want = _reference_flow(wb, *j, truncate_percent=None)
        assert rel_err(g, want) < 1e-4 and abs(float(g.sum()) - 1.0) < 1e-3
    rows = IG.mean_encodings(wb, [jobs[0][0]])
    assert rows.shape == (1, 256)


def _check_cubic_zoom(gpu):
    """the batched resize (xfrb_cubic_zoom) against scipy.ndimage.zoom(order=3, mode='grid-constant',
    grid_mode=True) + clip, which is how scikit-image >= 0.19 evaluates show.processSaliency's resize(order=3, mode='constant'),
    and against the per-map host function process_saliency."""
    import scipy.ndimage
    wb = whitebox.Whitebox(_net(L1111, gpu))
    rng = np.random.RandomState(3)
    maps = np.stack([rng.rand(112, 112).astype(np.float32) ** 3, np.outer(np.hanning(112), np.hanning(112)).astype(np.float32) + 2.0,
                     np.zeros((112, 112), np.float32)])
    maps[2, 40:44, 60:70] = 1e-3                                   # sparse map: ringing below zero must be clipped away
    got = IG.process_saliency_batch(wb, maps, (224, 224))
    assert got.shape == (3, 224, 224) and got.dtype == np.float32
    for g, m in zip(got, maps):
        want = IG.process_saliency(np.zeros((224, 224, 3)), m)
        assert float(np.abs(g - want).max()) < 2e-6 and g.min() >= 0.0 and g.max() <= 1.0
        n = (m - m.min()) / (m.max() - m.min() + np.float32(1e-9))
        z = scipy.ndimage.zoom(n.astype(np.float64), 2, order=3, mode='grid-constant', cval=0.0, grid_mode=True)
        assert float(np.abs(g - np.clip(z, n.min(), n.max())).max()) < 2e-6
    big = IG.process_saliency_batch(wb, maps[:1, :96, :80], (144, 200))           # other sizes / non-integer zooms
    n = (maps[0, :96, :80] - maps[0, :96, :80].min()) / (maps[0, :96, :80].max() - maps[0, :96, :80].min() + np.float32(1e-9))
    z = scipy.ndimage.zoom(n.astype(np.float64), (144 / 96, 200 / 80), order=3, mode='grid-constant', cval=0.0, grid_mode=True)
    assert float(np.abs(big[0] - np.clip(z, n.min(), n.max())).max()) < 2e-6


def test_cubic_zoom_emulated():
    _check_cubic_zoom(False)


@pytest.mark.gpu
def test_cubic_zoom_gpu():
    _check_cubic_zoom(True)

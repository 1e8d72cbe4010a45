"""the callers either side of the hot path in the inpainting-game flow.

Reference: python/xfr/inpainting_game/generate_whitebox_saliency.py (job functions, lines 45-219; file-name scheme,
lines 307-404) and python/xfr/show.py:131-137, 196-221 (post-processing + the .npz the evaluation reads,
plot_inpainting_game.py:228).  The reference builds one Whitebox per job and runs batch 1; here the per-job functions
keep their names and arguments and a *_batch twin feeds N jobs through one sweep of the batched engine:

    all mate / non-mate images of all jobs -> ONE batched encode -> per-job mean, L2-normalise (lines 57-69)
    -> Whitebox.net.set_triplet_classifiers(rows) -> Whitebox.contrastive_ebp_batch(probes)

run_contrastive_triplet_ebp_sharded spreads the jobs over the ranks of a torchrun job (contiguous shards, one gather of maps).

File discovery (the CSV / directory walk of generate_wb_smaps, lines 222-290) and the PNG overlay stay host glue of the
caller: they are image I/O, not arithmetic.

Row 3, scoring (python/xfr/inpainting_game/inpainting_game.py:12-146): `classified_as_inpainted_twin` keeps the reference's
signature.  The host turns the saliency map into ONE float64 value map + K thresholds (mask_value_map: the seeded noise, the
sort and the cumulative sum of 50,176 doubles); the K (= 101) blends are built on the device by xfrb_twin_blends straight
into the forward sweep's NHWC input (the reference materialises 101 float64 masks and 122 MB of float64 blends on the
host), swept by the same conv kernels as the saliency path, and only the K x D embeddings come back.
classified_as_inpainted_twin_batch puts the blends of several jobs through full engine sweeps.
"""
import os

import numpy as np
import torch

CLASSIFIER_SCALE = 1.0 / 2500.0     # generate_whitebox_saliency.py:103-104


def shorten_subtree_mode(ebp_subtree_mode):
    """generate_whitebox_saliency.py:217-220"""
    if ebp_subtree_mode == 'affineonly_with_prior':
        return 'awp'
    return ebp_subtree_mode


def saliency_method_name(wb, method, ebp_ver, device_type, truncate_percent=None, topk=32, subtree_mode_weighted=None):
    """The `method` component of '<mask_id>-<method>-saliency.npz' (generate_whitebox_saliency.py:310-378)."""
    mode = shorten_subtree_mode(wb.ebp_subtree_mode())
    if method == 'meanEBP':
        return 'meanEBP_mode=%s_v%02d_%s' % (mode, ebp_ver, device_type)
    if method == 'contrastive':
        if truncate_percent is None:
            return 'contrastive_triplet_ebp_mode=%s_v%02d_%s' % (mode, ebp_ver, device_type)
        return 'trunc_contrastive_triplet_ebp_mode=%s_v%02d_pct%d_%s' % (mode, ebp_ver, truncate_percent, device_type)
    if method == 'weighted-subtree':
        return 'weighted_subtree_triplet_ebp_mode=%s,%s_v%02d_top%d_%s' % (
            mode, shorten_subtree_mode(subtree_mode_weighted), ebp_ver, topk, device_type)
    raise RuntimeError("Unknown method type %s (valid types: 'meanEBP', 'contrastive', 'weighted-subtree')" % method)


def _to_device(wb, images, device):
    x = torch.cat([wb.convert_from_numpy(im) for im in images])
    return x if device is None else x.to(device)


def mean_encodings(wb, image_lists, device=None):
    """image_lists: J lists of HxWx3 images -> [J, D] unit-norm mean encodings (generate_whitebox_saliency.py:57-69:
    torch.mean over the stacked [1, D] encodings, then divide by torch.norm).  One batched encode for all lists."""
    counts = [len(l) for l in image_lists]
    if min(counts) == 0:
        raise ValueError('mean_encodings: an empty image list (the reference would take the mean of an empty stack)')
    flat = [im for l in image_lists for im in l]
    enc = wb.encode(_to_device(wb, flat, device)).detach()
    out, o = [], 0
    for c in counts:
        avg = torch.mean(enc[o:o + c], dim=0)
        out.append(avg / torch.norm(avg))
        o += c
    return torch.stack(out)


def run_contrastive_triplet_ebp_batch(wb, jobs, truncate_percent=None, device=None):
    """jobs: N tuples (im_mates, im_nonmates, probe_im).  Returns [N, h, w] maps, map i being what the reference's
    run_contrastive_triplet_ebp (lines 81-118) returns for job i."""
    x_mate = mean_encodings(wb, [j[0] for j in jobs], device)
    x_nonmate = mean_encodings(wb, [j[1] for j in jobs], device)
    wb.net.set_triplet_classifiers(CLASSIFIER_SCALE * x_mate, CLASSIFIER_SCALE * x_nonmate)
    probes = _to_device(wb, [j[2] for j in jobs], device)
    return wb.contrastive_ebp_batch(probes, 0, 1, percentile=truncate_percent)


def run_contrastive_triplet_ebp_sharded(wb, jobs, truncate_percent=None, device=None, dst=0):
    """Multi-GPU form of the batch (the reference farms jobs out one process per GPU,
    eval/generate_inpaintinggame_wb_saliency_maps_multigpu.py:191-231): one process per GPU under torchrun, rank r sweeps
    the contiguous slice shard_range(len(jobs), r, world) of the jobs and the only collective is the gather of the finished
    maps to rank `dst`.  Returns the [N, h, w] maps in job order on `dst`, None on the other ranks."""
    import torch.distributed as dist
    from .shard import gather_maps, shard_range
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return run_contrastive_triplet_ebp_batch(wb, jobs, truncate_percent, device)
    lo, hi = shard_range(len(jobs), dist.get_rank(), dist.get_world_size())
    hw = wb.net.engine().map_hw
    if hi > lo:
        local = torch.from_numpy(np.ascontiguousarray(run_contrastive_triplet_ebp_batch(wb, jobs[lo:hi], truncate_percent, device)))
    else:
        local = torch.empty(0, hw, hw, dtype=torch.uint8 if wb.convert_saliency_uint8 else torch.float32)
    if dist.get_backend() == 'nccl':
        local = local.to(wb.net._device())
    out = gather_maps(local, len(jobs), dst=dst)
    return None if out is None else out.cpu().numpy()


def run_contrastive_triplet_ebp(wb, im_mates, im_nonmates, probe_im, net_name=None, ebp_version=None, truncate_percent=None,
                                device=None):
    """generate_whitebox_saliency.py:81-118 (same arguments; net_name / ebp_version are unused there too)."""
    return run_contrastive_triplet_ebp_batch(wb, [(im_mates, im_nonmates, probe_im)], truncate_percent, device)[0]


# ebp_version -> (do_max_subtree, do_mated_similarity_gating), generate_whitebox_saliency.py:167-190; other versions: (False, False)
_SUBTREE_FLAGS = {7: (True, True), 8: (False, True), 9: (True, False), 10: (True, True), 11: (True, True), 12: (False, True)}


def run_weighted_subtree_triplet_ebp(wb, im_mates, im_nonmates, probe_im, net_name=None, subtree_mode_weighted='norelu',
                                     ebp_version=None, device=None, topk=1):
    """generate_whitebox_saliency.py:122-205: unit-norm (un-scaled) classifier rows, flags from ebp_version."""
    rows = mean_encodings(wb, [im_mates, im_nonmates], device)
    wb.net.set_triplet_classifier(rows[0:1], rows[1:2])
    do_max_subtree, gating = _SUBTREE_FLAGS.get(ebp_version, (False, False))
    probe = _to_device(wb, [probe_im], device)
    img_subtree, P_img, P_subtree, k_subtree = wb.weighted_subtree_ebp(
        probe, k_poschannel=0, k_negchannel=1, topk=topk, verbose=False, do_max_subtree=do_max_subtree,
        subtree_mode=subtree_mode_weighted, do_mated_similarity_gating=gating)
    return img_subtree


def run_weighted_subtree_triplet_ebp_sharded(wb, jobs, subtree_mode_weighted='all', ebp_version=None, device=None, topk=32, dst=0):
    """BASELINE configs[3] (the inpainting-game set, eval/generate_inpaintinggame_wb_saliency_maps_multigpu.py:121-231 farms one
    job per process per GPU): `jobs` = [(im_mates, im_nonmates, probe_im), ...] are sharded contiguously over the ranks
    (one process per GPU under torchrun), every rank runs run_weighted_subtree_triplet_ebp on its slice - the sub-trees of one
    probe are already batched as gradient rows inside weighted_subtree_ebp - and the only collective is the gather of the
    finished maps to rank `dst`.  Returns the [len(jobs), h, w] maps in job order on `dst`, None elsewhere."""
    import torch.distributed as dist
    from .shard import gather_maps, shard_range
    multi = dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1
    lo, hi = shard_range(len(jobs), dist.get_rank(), dist.get_world_size()) if multi else (0, len(jobs))
    hw = wb.net.engine().map_hw
    maps = [run_weighted_subtree_triplet_ebp(wb, *job, subtree_mode_weighted=subtree_mode_weighted, ebp_version=ebp_version,
                                             device=device, topk=topk) for job in jobs[lo:hi]]
    local = np.stack(maps) if maps else np.empty((0, hw, hw), dtype=np.uint8 if wb.convert_saliency_uint8 else np.float32)
    if not multi:
        return local
    local = torch.from_numpy(np.ascontiguousarray(local))
    if dist.get_backend() == 'nccl':
        local = local.to(wb.net._device())
    out = gather_maps(local, len(jobs), dst=dst)
    return None if out is None else out.cpu().numpy()


def mean_ebp(wb, probe_im, net_name=None, ebp_version=None, device=None):
    """generate_whitebox_saliency.py:207-215: EBP with a uniform prior over the network's classes."""
    x = _to_device(wb, [probe_im], device)
    P = torch.ones((1, wb.net.num_classes()))
    return wb.ebp(x, P if device is None else P.to(device))


def process_saliency_batch(wb, maps, out_hw=(224, 224)):
    """show.processSaliency (show.py:131-137) for a batch of maps on the B200: min-shift, max-normalise and the cubic resize to
    the probe size as ONE kernel launch (xfrb_cubic_zoom: one CTA per map, spline coefficients in shared memory).  maps
    [N,h,w] float32 (numpy, or a device tensor straight from the sweep) -> numpy float32 [N,oh,ow].  The definition is the one of
    process_saliency below - scikit-image >= 0.19's resize(order=3, mode='constant') = scipy.ndimage.zoom(order=3,
    mode='grid-constant', grid_mode=True) + clip - and the kernel is pinned to scipy's result (tests/test_inpaintgame.py)."""
    be = wb.net.engine(wb._ebp_with_bias).be
    dev = wb.net._device()
    t = torch.as_tensor(np.ascontiguousarray(maps) if isinstance(maps, np.ndarray) else maps, dtype=torch.float32).to(dev).contiguous()
    out = torch.empty((t.shape[0],) + tuple(out_hw), dtype=torch.float32, device=dev)
    be.cubic_zoom(t, out, normalize=True)
    return out.cpu().numpy()


def process_saliency(img, attMap):
    """show.py:131-137: min-shift, max-normalise, cubic resize to the image size.  The resize is scikit-image's
    (README.md:37 pins >= 0.17.2; absent here): from 0.19 on `resize(order=3, mode='constant')` without anti-aliasing (an
    upscale) is scipy.ndimage.zoom(order=3, mode='grid-constant', grid_mode=True) followed by a clip to the input range
    (skimage/transform/_warps.py resize(): zoom_factors = 1 / factors, ndi.zoom(..., mode=ndi_mode, cval=cval, grid_mode=True),
    then _clip_warp_output), which is what runs here.  Parity against scikit-image itself is unpinned (it is not installed and no
    reference output exists); the definition is pinned to scipy, which is what scikit-image calls."""
    import scipy.ndimage
    attMap = np.asarray(attMap)
    attMap = attMap - attMap.min()
    attMap = attMap / (attMap.max() + 1e-9)
    oh, ow = np.asarray(img).shape[:2]
    if attMap.shape == (oh, ow):
        return attMap
    zoom = (oh / attMap.shape[0], ow / attMap.shape[1])
    out = scipy.ndimage.zoom(attMap, zoom, order=3, mode='grid-constant', cval=0.0, grid_mode=True)
    return np.clip(out, attMap.min(), attMap.max())


def save_smap(npz_filename, smap, probe_im):
    """The arithmetic of show.create_save_smap (show.py:212-220) without the PNG overlay: float32, min-shift,
    sum-normalise, process_saliency, np.savez_compressed(saliency_map=...).  Returns the stored array."""
    smap = np.array(smap, dtype=np.float32)
    smap -= smap.min()
    smap /= smap.sum()
    smap = process_saliency(probe_im, smap)
    d = os.path.dirname(npz_filename)
    if d:
        os.makedirs(d, exist_ok=True)
    np.savez_compressed(npz_filename, saliency_map=smap)
    return smap


def generate_wb_smaps_batch(wb, jobs, output_dirs, mask_ids, ebp_ver, device=None, method='contrastive', overwrite=True):
    """Batched counterpart of generate_wb_smaps' `contrastive` branch (generate_whitebox_saliency.py:329-363): for N jobs
    writes '<output_dir>/<mask_id>-<method name>-saliency.npz' for the plain and the 20-percent truncated map, two engine
    sweeps in total.  jobs as in run_contrastive_triplet_ebp_batch; returns the list of files written."""
    if method != 'contrastive':
        raise RuntimeError("generate_wb_smaps_batch handles method 'contrastive'; run 'meanEBP' / 'weighted-subtree' per job")
    dev_type = 'cpu' if device is None else torch.device(device).type
    written = []
    for pct in (None, 20):
        name = saliency_method_name(wb, 'contrastive', ebp_ver, dev_type, truncate_percent=pct)
        files = ['%s/%s-%s-saliency.npz' % (d, m, name) for d, m in zip(output_dirs, mask_ids)]
        todo = [i for i, f in enumerate(files) if overwrite or not os.path.exists(f)]
        if not todo:
            continue
        maps = run_contrastive_triplet_ebp_batch(wb, [jobs[i] for i in todo], pct, device)
        sizes = {np.asarray(jobs[i][2]).shape[:2] for i in todo}
        if len(sizes) == 1 and getattr(wb.net.engine(wb._ebp_with_bias).be, 'name', '') == 'cuda':
            # every probe of the game is 224x224: the .npz arithmetic of show.create_save_smap (show.py:212-220) for the whole batch,
            # the cubic resize as one kernel launch
            m32 = np.array(maps, dtype=np.float32)
            m32 -= m32.min(axis=(1, 2), keepdims=True)
            m32 /= m32.sum(axis=(1, 2), keepdims=True)
            big = process_saliency_batch(wb, m32, next(iter(sizes)))
            for i, m in zip(todo, big):
                d = os.path.dirname(files[i])
                if d:
                    os.makedirs(d, exist_ok=True)
                np.savez_compressed(files[i], saliency_map=m)
                written.append(files[i])
            continue
        for i, m in zip(todo, maps):
            save_smap(files[i], m, jobs[i][2])
            written.append(files[i])
    return written


# ------------------------------------------------------------------ row 3: inpainting-game scoring
def mask_value_map(saliency_map, threshold_method, percentiles=None, thresholds=None, seed=None, max_noise=1e-9,
                   include_zero_elements=True):
    """-> (value [H,W] float64, thr [K] float64) with the reference's masks == value[None] > thr[:, None, None]
    (inpainting_game.py:26-66).  The tie-breaking noise is np.random.seed(seed); np.random.rand(H, W) - drawn here from a
    private RandomState(seed), the same stream without touching the global generator."""
    saliency_map = np.asarray(saliency_map)
    keep = 1 if include_zero_elements else (saliency_map != 0)
    value = saliency_map + keep * np.random.RandomState(seed).rand(*saliency_map.shape) * max_noise
    value = value / value.sum()
    if threshold_method == 'percent-density':
        # every element is replaced by the cumulative mass of the elements not larger than it, in units of the total mass
        order = np.argsort(value.flat)
        cdf = np.cumsum(value.flat[order])
        value.flat[order] = cdf
        value = value / value.max()
        thr = 1.0 - np.asarray(percentiles).astype(value.dtype) / 100
        if percentiles[-1] == 100:
            thr[-1] = 0
    elif thresholds is None:
        thr = np.percentile(value, 100 - np.asarray(percentiles))
        if percentiles[0] == 0:
            thr[0] = 1
        if percentiles[-1] == 100:
            thr[-1] = 0
    else:
        thr = np.asarray(thresholds)
    return np.ascontiguousarray(value, dtype=np.float64), np.ascontiguousarray(thr, dtype=np.float64)


def create_threshold_masks(saliency_map, threshold_method, percentiles=None, thresholds=None, seed=None, max_noise=1e-9,
                           include_zero_elements=True, blur_sigma=None):
    """inpainting_game.py:12-80 (same arguments): [K,H,W] masks, True = take the inpainted pixel.  With blur_sigma the masks
    are floats of the saliency map's dtype, blurred with skimage.filters.gaussian (== scipy.ndimage.gaussian_filter,
    mode 'nearest', truncate 4) at sigma = blur_sigma percent of the shorter side; the 100 % mask stays hard."""
    value, thr = mask_value_map(saliency_map, threshold_method, percentiles, thresholds, seed, max_noise, include_zero_elements)
    masks = value[np.newaxis] > thr[:, np.newaxis, np.newaxis]
    if blur_sigma is not None and blur_sigma > 0:
        import scipy.ndimage
        masks = masks.astype(np.asarray(saliency_map).dtype)
        sigma = blur_sigma * np.min(np.asarray(saliency_map).shape) / 100.0
        for i in range(masks.shape[0]):
            if percentiles[i] != 100:
                masks[i] = scipy.ndimage.gaussian_filter(masks[i], sigma, mode='nearest', truncate=4.0)
    return masks


def _network_format(imT):
    a = imT.detach().cpu().numpy() if isinstance(imT, torch.Tensor) else np.asarray(imT)
    if a.ndim != 3 or a.shape[0] not in (1, 3):
        raise ValueError('classified_as_inpainted_twin: images must be in network format [C,H,W] with C in (1, 3), got %s '
                         '(Whitebox.embeddings asserts the same, whitebox.py:764)' % (a.shape,))
    return np.ascontiguousarray(a, dtype=np.float64)         # inpainting_game.py:124-125


def _twin_blends_into(snet, blends, original_imT, inpaint_imT, value=None, thr=None, masks=None, mask_f32=False):
    """Enqueue xfrb_twin_blends for one job into `blends` [K,H,W,C] (a slice of the sweep's input buffer)."""
    net = snet.net
    dev = net._device()
    o, p = _network_format(original_imT), _network_format(inpaint_imT)
    if o.shape != p.shape:
        raise ValueError('original / inpainted image shapes differ: %s vs %s' % (o.shape, p.shape))
    if tuple(blends.shape[1:]) != (o.shape[1], o.shape[2], o.shape[0]):
        raise ValueError('images of one batch must share their shape: %s vs blends %s' % (o.shape, tuple(blends.shape)))
    up = lambda a: None if a is None else torch.from_numpy(np.ascontiguousarray(a, dtype=np.float64)).to(dev)
    net.engine().be.twin_blends(up(o), up(p), up(value), up(thr), up(masks), blends, mask_f32)


def _normalised(embeds):
    flat = embeds.reshape((embeds.shape[0], -1))
    return (flat / np.linalg.norm(flat, axis=1, keepdims=True)).reshape(embeds.shape)


def twin_blend_embeddings(snet, original_imT, inpaint_imT, value=None, thr=None, masks=None, mask_f32=False):
    """Device side of the scoring: K blends -> forward sweep -> [K, D] embeddings, normalised as Whitebox.embeddings does
    (whitebox.py:779-783).  snet: xfr_b200.whitebox.Whitebox."""
    C, H, W = _network_format(original_imT).shape
    K = (masks if masks is not None else thr).shape[0]
    blends = torch.empty(K, H, W, C, device=snet.net._device(), dtype=torch.float32)
    _twin_blends_into(snet, blends, original_imT, inpaint_imT, value, thr, masks, mask_f32)
    return _normalised(snet.net.encode_nhwc(blends).detach().cpu().numpy())


def classified_as_inpainted_twin(snet, original_imT, inpaint_imT, original_gal_embed, inpaint_gal_embed, saliency_map,
                                 mask_threshold_method, include_zero_elements=True, mask_blur_sigma=None, percentiles=None,
                                 thresholds=None, seed=None, binary_classification=True, return_transitions=False):
    """inpainting_game.py:83-146 (same arguments and returns): replace the most salient pixels of the original probe by the
    inpainted twin's, percentile by percentile, and report which blends the matcher places closer to the twin's gallery."""
    value, thr = mask_value_map(saliency_map, mask_threshold_method, percentiles, thresholds, seed,
                                include_zero_elements=include_zero_elements)
    masks, mask_f32 = None, False
    if mask_blur_sigma is not None and mask_blur_sigma > 0:
        masks = create_threshold_masks(saliency_map, mask_threshold_method, percentiles, thresholds, seed,
                                       include_zero_elements=include_zero_elements, blur_sigma=mask_blur_sigma)
        mask_f32 = masks.dtype == np.float32         # numpy then rounds `1.0 - masks` to float32 (inpainting_game.py:128)
        if masks.dtype not in (np.float32, np.float64):
            raise ValueError('blurred masks need a float32 / float64 saliency map, got %s' % masks.dtype)
    blend_embeds = twin_blend_embeddings(snet, original_imT, inpaint_imT, value, thr, masks, mask_f32)
    blend_embeds = blend_embeds / np.linalg.norm(blend_embeds, axis=1, keepdims=True)
    pr_dist = np.linalg.norm(blend_embeds - original_gal_embed, axis=1)
    pg_dist = np.linalg.norm(blend_embeds - inpaint_gal_embed, axis=1)
    classified_as_twin = pg_dist < pr_dist
    assert not classified_as_twin[0]
    if not return_transitions:
        return classified_as_twin, pg_dist, pr_dist
    # the blends and masks themselves are only returned for plotting: materialised on the host, as the reference does
    if masks is None:
        masks = value[np.newaxis] > thr[:, np.newaxis, np.newaxis]
    o, p = _network_format(original_imT), _network_format(inpaint_imT)
    cm = masks[:, np.newaxis] if o.shape[0] == 1 else np.repeat(masks[:, np.newaxis], 3, axis=1)      # dtype kept: see mask_f32
    blends = (1.0 - cm) * o[np.newaxis] + cm * p[np.newaxis]
    return classified_as_twin, pg_dist, pr_dist, blends, masks


def classified_as_inpainted_twin_batch(snet, jobs, mask_threshold_method, include_zero_elements=True, percentiles=None,
                                       thresholds=None, seed=None):
    """Batched scoring: jobs = [(original_imT, inpaint_imT, original_gal_embed, inpaint_gal_embed, saliency_map), ...] with the
    arguments of classified_as_inpainted_twin (hard masks; `seed` one value or one per job).  The K blends of every job go into
    ONE device buffer and through the forward sweep together (full engine sweeps instead of one K-blend sweep per job); one
    device-to-host copy of the [J*K, D] embeddings.  Returns [(classified_as_twin, pg_dist, pr_dist), ...] in job order."""
    if not jobs:
        return []
    seeds = list(seed) if isinstance(seed, (list, tuple, np.ndarray)) else [seed] * len(jobs)
    maps = [mask_value_map(j[4], mask_threshold_method, percentiles, thresholds, sd, include_zero_elements=include_zero_elements)
            for j, sd in zip(jobs, seeds)]
    K = maps[0][1].shape[0]
    C, H, W = _network_format(jobs[0][0]).shape
    blends = torch.empty(len(jobs) * K, H, W, C, device=snet.net._device(), dtype=torch.float32)
    for i, (j, (value, thr)) in enumerate(zip(jobs, maps)):
        _twin_blends_into(snet, blends[i * K:(i + 1) * K], j[0], j[1], value, thr)
    embeds = snet.net.encode_nhwc(blends).detach().cpu().numpy()
    out = []
    for i, j in enumerate(jobs):
        e = _normalised(embeds[i * K:(i + 1) * K])
        e = e / np.linalg.norm(e, axis=1, keepdims=True)
        pr_dist = np.linalg.norm(e - j[2], axis=1)
        pg_dist = np.linalg.norm(e - j[3], axis=1)
        cls = pg_dist < pr_dist
        assert not cls[0]
        out.append((cls, pg_dist, pr_dist))
    return out

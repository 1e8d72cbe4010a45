"""the drop-in boundary.  The overlay package dropin/xfr puts the B200 engine behind the reference's own module name
`xfr.models.whitebox`; tests/dropin_driver.py then runs the reference's UNMODIFIED caller code (generate_whitebox_saliency.py job
functions, inpainting_game.py scoring, the demo's call sequence on a CPU-resident network) on top of this repo's classes and
compares with the reference-generated goldens.  Needs /root/reference (build container); the kernels there are the emulation."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = '/root/reference/python'


def _run_driver(backend):
    env = dict(os.environ)
    env['PYTHONPATH'] = os.pathsep.join([os.path.join(ROOT, 'dropin'), os.path.join(ROOT, 'oracle', 'shim'), REF])
    p = subprocess.run([sys.executable, os.path.join(ROOT, 'tests', 'dropin_driver.py'), '--backend', backend], env=env,
                       capture_output=True, text=True, timeout=1500)
    assert p.returncode == 0, p.stderr[-3000:]
    return json.loads(p.stdout.strip().splitlines()[-1])


def _check(out, tol_c, tol):
    assert out['whitebox_module'].endswith(os.path.join('dropin', 'xfr', 'models', 'whitebox.py'))
    assert out['xfr_root'] == '/root/reference' and out['refgen'].startswith('/root/reference/')
    e = out['err']
    for k, (a, r) in e.items():
        if k.startswith('job') or k in ('demo_cebp_awp_smooth', 'demo_tcebp20_awp_smooth', 'demo_cebp_fc2head'):
            assert a < 1e-4 and r < tol_c, (k, a, r)                   # contrastive maps: the 1e-4 max-abs bar + scale-aware
        elif k in ('ws_eval_smap', 'ws_v7_smap', 'mean_ebp', 'demo_ebp_fc2head', 'demo_ebp_awp_smooth', 'demo_encode', 'twin_pg_dist'):
            assert r < tol, (k, a, r)
    assert e['twin_cls_mismatches'][0] == 0
    assert e['demo_ws_u8_pixels_off_by_more_than_2'][0] <= 5 and e['demo_ws_k_set_diff'][0] <= 2
    assert out['P_len'] == 59 and out['P_names_ok'] and out['P_sums_rel'] < tol
    assert out['layerlist'] > 50


@pytest.mark.skipif(not os.path.isdir(REF), reason='needs the reference tree (build container only)')
def test_reference_callers_over_dropin_emulated():
    _check(_run_driver('emul'), 1e-3, 2e-3)


def test_overlay_needs_the_reference_behind_it():
    """Without the reference package behind it on sys.path the overlay says so (ImportError), it does not half-import."""
    env = dict(os.environ)
    env['PYTHONPATH'] = os.path.join(ROOT, 'dropin')
    p = subprocess.run([sys.executable, '-c', 'import xfr'], env=env, capture_output=True, text=True, timeout=300, cwd='/tmp')
    assert p.returncode != 0 and 'drop-in overlay' in p.stderr


@pytest.mark.gpu
@pytest.mark.skipif(not os.path.isdir(REF), reason='needs the reference tree, which does not travel to the GPU box')
def test_reference_callers_over_dropin_gpu():
    _check(_run_driver('cuda'), 5e-2, 1e-2)

"""GPU: the conv-kernel forms that are selected by environment switches must agree with the default selection.

The switches are read once per process, so every variant runs in a subprocess that prints a digest of one bf16 U-Net forward
at a batch where the persistent launches (and with them halo mode / the CTA-pair kernel) are taken:

  * `SPDM_PAIR=1`  -- the `cta_group::2` CTA-pair form of the 256-wide kernel (off by default: measured slower).  Same MMAs in the
    same K order, fp32 accumulation in tensor memory: the output must match the single-CTA form to bf16 rounding of a few outputs.
  * `SPDM_NO_HALO=1`, `SPDM_NO_PIX256=1`, `SPDM_FUSE_EPI16=0`, `SPDM_FUSE_SHORT=0` -- the operand-staging / epilogue variants the
    defaults replaced this round (different tap order or epilogue split: fp32 summation order changes, bf16-level agreement).
"""
import json
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

SCRIPT = r'''
import json, sys, torch
sys.path.insert(0, %r)
from oracle import fixtures
import state_policy_diffusionmodel_b200 as spdm
import os
B = int(os.environ.get("SPDM_TEST_B", "1024"))
ROWS = int(os.environ.get("SPDM_TEST_ROWS", "31"))
sd = fixtures.make_unet_weights(attention=True, seed=0)
plan = spdm.DenoisePlan(attention=True, precision="bf16", batch_max=B, inpaint_rows=1, rows=ROWS, dim=5)
plan.load_unet_state_dict(sd)
g = torch.Generator().manual_seed(5)
x = torch.randn((B, 1, ROWS, 5), generator=g).cuda()
t = torch.randint(0, 1000, (B,), generator=g).cuda()
y = (0.5 * torch.randn((B, 1350), generator=g)).cuda()
eps = plan.unet_forward(x, t, y).float().cpu()
print("DIGEST " + json.dumps({"sum": float(eps.double().sum()), "abs": float(eps.double().abs().sum()),
                              "head": [float(v) for v in eps.flatten()[:64]], "max": float(eps.abs().max())}))
''' % ROOT


def _run(env_extra):
    env = dict(os.environ)
    for k in ("SPDM_PAIR", "SPDM_NO_HALO", "SPDM_NO_PIX256", "SPDM_FUSE_EPI16", "SPDM_FUSE_SHORT", "SPDM_TEST_B", "SPDM_TEST_ROWS"):
        env.pop(k, None)
    env.update(env_extra)
    res = subprocess.run([sys.executable, "-c", SCRIPT], capture_output=True, text=True, env=env, timeout=600)
    assert res.returncode == 0, res.stdout[-2000:] + res.stderr[-2000:]
    line = [l for l in res.stdout.splitlines() if l.startswith("DIGEST ")][-1]
    return json.loads(line[len("DIGEST "):])


@pytest.fixture(scope="module")
def default_digest():
    return _run({})


def _close(a, b, tol):
    scale = max(a["max"], 1e-6)
    worst = max(abs(x - y) for x, y in zip(a["head"], b["head"])) / scale
    assert worst <= tol, worst
    assert abs(a["abs"] - b["abs"]) / a["abs"] <= tol, (a["abs"], b["abs"])


def test_cta_pair_kernel_matches_single_cta_form(default_digest):
    _close(default_digest, _run({"SPDM_PAIR": "1"}), 2e-3)


@pytest.mark.parametrize("env", [{"SPDM_NO_HALO": "1"}, {"SPDM_NO_PIX256": "1"}, {"SPDM_FUSE_EPI16": "0", "SPDM_FUSE_SHORT": "0"}])
def test_operand_staging_and_epilogue_variants_match_defaults(default_digest, env):
    _close(default_digest, _run(env), 5e-3)


@pytest.mark.parametrize("rows,B", [(61, 256), (121, 128)])
def test_halo_mode_on_the_longer_horizons(rows, B):
    """64x8 / 128x8 maps: a 256-pixel tile is 32 rows in the MIDDLE of a sample, so the halo rows are real neighbours (not the
    out-of-bounds zero fill of the 32x8 map); persistent launches (more tiles than SMs), halo on vs off."""
    geo = {"SPDM_TEST_ROWS": str(rows), "SPDM_TEST_B": str(B)}
    _close(_run(geo), _run(dict(geo, SPDM_NO_HALO="1")), 5e-3)

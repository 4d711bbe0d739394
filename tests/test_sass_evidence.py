"""The built library really contains the Blackwell data path it claims (B200_PROFILING.md: "SASS mnemonics that prove TMA"):
k_tile3d must carry TMA tensor loads (UTMALDG), TMA bulk loads (UBLKCP), mbarrier operations (SYNCS), packed f32x2 arithmetic
(FFMA2 / FADD2), warp match (MATCH) for the group reduction and, in its J-tile variant, the TMA reduce (UTMAREDG).  Needs only
cuobjdump (no GPU)."""
import os
import re
import shutil
import subprocess

import pytest

from pypic3d_b200 import _lib


def _cuobjdump():
    for c in ("/usr/local/cuda/bin/cuobjdump", shutil.which("cuobjdump")):
        if c and os.path.exists(c):
            return c
    return None


@pytest.fixture(scope="module")
def tile_kernels():
    tool = _cuobjdump()
    if tool is None:
        pytest.skip("cuobjdump not available")
    path = _lib.build(force=False)
    out = subprocess.run([tool, "-sass", path], capture_output=True, text=True, check=True).stdout
    kernels, name = {}, None
    for line in out.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            name = m.group(1)
            kernels.setdefault(name, [])
        elif name is not None:
            kernels[name].append(line)
    tile = {k: "\n".join(v) for k, v in kernels.items() if "k_tile3d" in k}
    assert tile, "no k_tile3d kernels in the library"
    return tile


def test_tile_kernel_uses_tma_mbarriers_and_packed_math(tile_kernels):
    f32 = {k: v for k, v in tile_kernels.items() if "k_tile3dIf" in k}
    assert len(f32) >= 12           # 2 pushers x 2 move variants x 3 reduction modes
    for name, sass in f32.items():
        assert sass.count("UTMALDG.3D") >= 6, name          # six E/B component boxes per stage request
        assert "UBLKCP" in sass and "SYNCS" in sass, name   # particle slices by TMA bulk copy; mbarrier waits / arrives
        assert "FFMA2" in sass and "FADD2" in sass, name    # f32x2 gather interpolation and reduction adds


def test_reduction_variants_are_the_ones_described(tile_kernels):
    # template argument 6 (MODE): 0 segmented scan, 1 shared-memory J tile + TMA reduce, 2 match-any groups
    mode = lambda k: re.search(r"Lb[01]ELi([012])EEEv", k).group(1)
    f32 = {k: v for k, v in tile_kernels.items() if "k_tile3dIf" in k}
    for name, sass in f32.items():
        m = mode(name)
        assert ("UTMAREDG" in sass) == (m == "1"), name
        assert ("MATCH.ANY" in sass) == (m == "2"), name
        # every variant keeps global fp32 adds for out-of-tile particles.  MODE 0/2 must issue them fire-and-forget (REDG);
        # the J-tile build currently gets the returning form (ATOMG) because its __threadfence_block() calls make ptxas
        # promote every RED in the kernel -- profiles/r01_k1_versions.md "J tile" records this as a confound of that A/B.
        if m == "1":
            assert "REDG.E.ADD.F32" in sass or "ATOMG.E.ADD.F32" in sass, name
            assert "ATOMS.CAST.SPIN" in sass, name          # shared fp32 adds are CAS loops on sm_100a
        else:
            assert "REDG.E.ADD.F32" in sass and "ATOMG.E.ADD.F32" not in sass, name


@pytest.fixture(scope="module")
def pair_kernels():
    tool = _cuobjdump()
    if tool is None:
        pytest.skip("cuobjdump not available")
    path = _lib.build(force=False)
    out = subprocess.run([tool, "-sass", path], capture_output=True, text=True, check=True).stdout
    kernels, name = {}, None
    for line in out.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            name = m.group(1)
            kernels.setdefault(name, [])
        elif name is not None:
            kernels[name].append(line)
    return {k: "\n".join(v) for k, v in kernels.items() if "k_pair3d" in k or "k_pair_fixup" in k or "k_yee_stream" in k or "k_boxes" in k}


def test_pair_kernel_is_tma_fed_packed_and_fire_and_forget(pair_kernels):
    """K1 v10: the benchmarked kernel.  TMA box + bulk loads and mbarriers, packed f32x2 arithmetic incl. the round-down add of
    the magic-constant floor (FADD2.RM), and -- the regression ncu caught in round 2 -- every deposit add a fire-and-forget REDG:
    one __threadfence_block() anywhere in the kernel makes ptxas promote them all to returning ATOMG (+0.5 ms per launch)."""
    f32 = {k: v for k, v in pair_kernels.items() if "k_pair3dIf" in k}
    assert len(f32) >= 12           # 2 pushers x 2 move variants x 3 reduction modes
    for name, sass in f32.items():
        assert sass.count("UTMALDG.3D") >= 6 and "UBLKCP" in sass and "SYNCS" in sass, name
        assert "FFMA2" in sass and "FADD2" in sass and "FADD2.RM" in sass, name
        assert "REDG.E.ADD.F32" in sass and "ATOMG" not in sass, name
        assert "LDL" not in sass and "STL" not in sass, name                 # no spills in the hot kernel
    f64 = {k: v for k, v in pair_kernels.items() if "k_pair3dId" in k}
    for name, sass in f64.items():
        assert "REDG.E.ADD.F64" in sass and "ATOMG" not in sass, name


def test_fixup_and_field_kernels_exist_in_both_dtypes(pair_kernels):
    for stem in ("k_pair_fixupIf", "k_pair_fixupId", "k_yee_streamIf", "k_yee_streamId", "k_boxesIfLi0", "k_boxesIfLi1", "k_boxesIdLi1"):
        assert any(stem in k for k in pair_kernels), stem
    for name, sass in pair_kernels.items():
        if "k_pair_fixup" in name:
            assert "REDG.E.ADD" in sass, name                                # union-stencil deposit of the cell changers
        if "k_boxes" in name and "Li1E" in name:
            assert "RED" in sass or "ATOM" in sass, name                     # accumulating unpack is atomic (overlapping boxes)

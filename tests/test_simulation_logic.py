"""Host-side decisions of the resident path that need no device: which K1 variant / reduction a launch gets."""
from types import SimpleNamespace

import pytest

from pypic3d_b200.simulation import Simulation


def _bare(**kw):
    sim = Simulation.__new__(Simulation)
    base = dict(p=SimpleNamespace(dx=1.0, dy=2.0, dz=0.5, dt=0.1, shape_factor=1, g=2, pusher=1, tile=(8, 8, 4), gmesh=(1, 1, 1)),
                deposition=0, ext_E=None, sort_interval=10, step_count=0, _vrms=[1.0, 0.01], _sorted_at=[0, 0],
                _jtile_mode="0", _groupred_mode="auto", _red_mode="scan", k1_variant="tile")
    base.update(kw)
    for k, v in base.items():
        setattr(sim, k, v)
    return sim


def test_k1_variant_selection(monkeypatch):
    monkeypatch.delenv("PIC_K1_VARIANT", raising=False)
    assert _bare()._pick_k1_variant(None) == "pair"            # K1 v10 is the default for the headline configuration
    monkeypatch.setenv("PIC_K1_VARIANT", "tile")
    assert _bare()._pick_k1_variant(None) == "tile"            # K1 v9, the A/B control
    monkeypatch.delenv("PIC_K1_VARIANT", raising=False)
    for bad in (dict(deposition=1), dict(ext_E=[object()]),
                dict(p=SimpleNamespace(shape_factor=2, g=2, pusher=1, tile=(8, 8, 4), gmesh=(1, 1, 1))),
                dict(p=SimpleNamespace(shape_factor=1, g=1, pusher=1, tile=(8, 8, 4), gmesh=(1, 1, 1))),
                dict(p=SimpleNamespace(shape_factor=1, g=2, pusher=2, tile=(8, 8, 4), gmesh=(1, 1, 1))),      # Higuera-Cary
                dict(p=SimpleNamespace(shape_factor=1, g=2, pusher=1, tile=(8, 6, 4), gmesh=(1, 1, 1))),      # width not a multiple of 4
                dict(p=SimpleNamespace(shape_factor=1, g=2, pusher=1, tile=(8, 8, 1), gmesh=(1, 1, 1)))):     # reduced axis
        assert _bare(**bad)._pick_k1_variant(None) == "global", bad
    monkeypatch.setenv("PIC_K1_VARIANT", "global")
    assert _bare()._pick_k1_variant(None) == "global"


def test_group_reduction_is_chosen_by_drift_since_the_last_sort():
    sim = _bare()
    # drift = vrms * dt * (steps since sort) / min(d) ; threshold 4 % of a cell: species 0 (0.2 cells per step) crosses it at once,
    # species 1 (0.002 per step) after 21 steps
    assert sim._k1_options(0) == 0 and sim._k1_options(1) == 0
    sim.step_count = 1
    assert sim._k1_options(0) == 2 and sim._k1_options(1) == 0
    sim.step_count = 21
    assert sim._k1_options(1) == 2
    sim._sorted_at = [21, 21]          # a sort resets the clock
    assert sim._k1_options(0) == 0 and sim._k1_options(1) == 0


@pytest.mark.parametrize("mode,expect", [("0", 0), ("1", 2)])
def test_group_reduction_can_be_forced(mode, expect):
    sim = _bare(_groupred_mode=mode, step_count=5)
    assert sim._k1_options(0) == expect and sim._k1_options(1) == expect
    assert _bare(_jtile_mode="1", _groupred_mode="0")._k1_options(0) == 1

"""oracle/metrics_oracle.py (the CPU restatement of the reference's validation metrics) against tests/golden/metrics.npz,
which oracle/make_golden_metrics.py produced by running the reference's own functions (Distribution.py:248-331,
Spatial_Distribution.py:362-493).  Bit-exact: same library, same operations, same order."""
import os

import numpy as np
import pytest

from oracle import metrics_oracle as MO


@pytest.fixture(scope="module")
def gold(golden_dir):
    return np.load(os.path.join(golden_dir, "metrics.npz"))


@pytest.mark.parametrize("tag", ["sce1", "sce4"])
def test_velocities_js_and_occupancy_equal_the_reference(gold, tag):
    gen, hum, name = gold[f"{tag}/gen"], gold[f"{tag}/hum"], str(gold[f"{tag}/model_name"])
    assert gen.dtype == np.float32 and gen.shape[1:] == (10, 3)
    vg, vh = MO.waypoint_velocities(list(gen)), MO.waypoint_velocities(list(hum))
    np.testing.assert_array_equal(vg, gold[f"{tag}/v_gen"])
    np.testing.assert_array_equal(vh, gold[f"{tag}/v_hum"])
    assert vg[0] == 0.0                                         # a repeated time stamp at the very beginning: nothing before it
    np.testing.assert_array_equal(MO.js_edges(vg, vh), gold[f"{tag}/bins_js"])
    hg, _ = np.histogram(vg, bins=gold[f"{tag}/bins_js"])
    np.testing.assert_array_equal(hg, gold[f"{tag}/hist_gen"])
    assert abs(MO.js_divergence(vg, vh) - float(gold[f"{tag}/js"])) < 1e-12
    Hg, Hh = MO.trajectories_per_cell(list(gen), name), MO.trajectories_per_cell(list(hum), name)
    np.testing.assert_array_equal(Hg, gold[f"{tag}/H_gen"])
    np.testing.assert_array_equal(Hh, gold[f"{tag}/H_hum"])
    xe, ye = MO.grid_edges(name)
    np.testing.assert_array_equal(xe, gold[f"{tag}/x_edges"])
    np.testing.assert_array_equal(ye, gold[f"{tag}/y_edges"])
    assert abs(MO.rmse_frequency(Hg, Hh) - float(gold[f"{tag}/rmse"])) < 1e-12
    assert Hg.max() <= len(gen) and Hg.sum() >= len(gen)      # a trajectory counts once per cell it visits

"""Logger-compatible export against the reference's own Logger output (tests/golden/logger.npz)."""
import glob
import os

import numpy as np

from helpers import load_golden
from gpd_b200.utils.Logger import Logger


def test_logger_matches_reference_formats(tmp_path):
    g = load_golden("logger.npz")
    states, controls, hz = g["in_states"], g["in_controls"], int(g["hz"])
    T, n = states.shape[0], states.shape[1]
    lg = Logger(logging_freq_hz=hz, output_folder=str(tmp_path), num_drones=n)
    for t in range(T):
        if t % 2:
            lg.log_batch(t / hz, states[t], controls[t])          # batched entry point
        else:
            for j in range(n):
                lg.log(drone=j, timestamp=t / hz, state=states[t, j], control=controls[t, j])
    assert np.array_equal(lg.states, g["states"]) and np.array_equal(lg.timestamps, g["timestamps"])
    assert np.array_equal(lg.controls, g["controls"])
    path = lg.save()
    z = np.load(path)
    assert sorted(z.files) == ["controls", "states", "timestamps"] and np.array_equal(z["states"], g["states"])
    csv_dir = lg.save_as_csv("kat")
    assert sorted(os.listdir(csv_dir)) == list(g["csv_names"])
    for key, text in zip(g["csv_keys"], g["csv_texts"]):
        assert open(os.path.join(csv_dir, str(key))).read() == str(text), key

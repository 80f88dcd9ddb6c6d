"""Trajectory logger writing the reference's on-disk formats (reference ``utils/Logger.py``): the in-memory layout
``states (N, 16, T)`` ordered ``[pos3, vel3, rpy3, ang_v3, rpm4]`` (``:117``), ``np.savez(timestamps, states,
controls)`` into a ``.npy`` file (``:123-128``) and the per-signal CSV files (``:131-203``).  Plotting (matplotlib)
is out of scope.  ``log_batch`` takes the (N, 20) state rows of one env straight from the batched simulator."""
from __future__ import annotations

import os
from datetime import datetime

import numpy as np

#: state20 -> Logger order (Logger.py:117): pos, vel, rpy, ang_v + rpm
_ORDER = np.r_[0:3, 10:13, 7:10, 13:20]
#: CSV file stem -> row of `states` (Logger.py:147-201)
_DIRECT = [("x", 0), ("y", 1), ("z", 2), ("r", 6), ("p", 7), ("ya", 8), ("vx", 3), ("vy", 4), ("vz", 5),
           ("wx", 9), ("wy", 10), ("wz", 11)]
_RATES = [("rr", 6), ("pr", 7), ("yar", 8)]


class Logger(object):
    def __init__(self, logging_freq_hz: int, output_folder: str = "results", num_drones: int = 1, duration_sec: int = 0,
                 colab: bool = False):
        self.COLAB = colab
        self.OUTPUT_FOLDER = output_folder
        os.makedirs(self.OUTPUT_FOLDER, exist_ok=True)
        self.LOGGING_FREQ_HZ = logging_freq_hz
        self.NUM_DRONES = num_drones
        self.PREALLOCATED_ARRAYS = duration_sec != 0
        n = duration_sec * logging_freq_hz
        self.counters = np.zeros(num_drones)
        self.timestamps = np.zeros((num_drones, n))
        self.states = np.zeros((num_drones, 16, n))
        self.controls = np.zeros((num_drones, 12, n))

    def _slot(self, drone):
        c = int(self.counters[drone])
        if c >= self.timestamps.shape[1]:
            self.timestamps = np.concatenate((self.timestamps, np.zeros((self.NUM_DRONES, 1))), axis=1)
            self.states = np.concatenate((self.states, np.zeros((self.NUM_DRONES, 16, 1))), axis=2)
            self.controls = np.concatenate((self.controls, np.zeros((self.NUM_DRONES, 12, 1))), axis=2)
        elif not self.PREALLOCATED_ARRAYS and self.timestamps.shape[1] > c:
            c = self.timestamps.shape[1] - 1
        return c

    def log(self, drone: int, timestamp, state, control=np.zeros(12)):
        """One step of one drone (Logger.py:83-121): ``state`` is the (20,) state vector, ``control`` (12,)."""
        state = np.asarray(state, dtype=np.float64)
        control = np.asarray(control, dtype=np.float64)
        if drone < 0 or drone >= self.NUM_DRONES or timestamp < 0 or len(state) != 20 or len(control) != 12:
            raise ValueError("[ERROR] in Logger.log(), invalid data")
        c = self._slot(drone)
        self.timestamps[drone, c] = timestamp
        self.states[drone, :, c] = state[_ORDER]
        self.controls[drone, :, c] = control
        self.counters[drone] = c + 1

    def log_batch(self, timestamp, states, controls=None):
        """One step of every drone of one env: ``states`` (N, 20) ndarray or tensor, ``controls`` (N, 12) or None."""
        if hasattr(states, "detach"):
            states = states.detach().double().cpu().numpy()
        if controls is not None and hasattr(controls, "detach"):
            controls = controls.detach().double().cpu().numpy()
        for j in range(self.NUM_DRONES):
            self.log(j, timestamp, states[j], np.zeros(12) if controls is None else controls[j])

    def save(self):
        """``np.savez(timestamps=, states=, controls=)`` into ``save-flight-<date>.npy`` (Logger.py:123-128)."""
        path = os.path.join(self.OUTPUT_FOLDER, "save-flight-" + datetime.now().strftime("%m.%d.%Y_%H.%M.%S") + ".npy")
        with open(path, 'wb') as out_file:
            np.savez(out_file, timestamps=self.timestamps, states=self.states, controls=self.controls)
        return path

    def save_as_csv(self, comment: str = ""):
        """Per-signal CSVs, file names and contents as Logger.py:131-203."""
        csv_dir = os.path.join(self.OUTPUT_FOLDER, "save-flight-" + comment + "-" + datetime.now().strftime("%m.%d.%Y_%H.%M.%S"))
        os.makedirs(csv_dir, exist_ok=True)
        t = np.arange(0, self.timestamps.shape[1] / self.LOGGING_FREQ_HZ, 1 / self.LOGGING_FREQ_HZ)

        def put(name, values):
            with open(os.path.join(csv_dir, name + ".csv"), 'wb') as out_file:
                np.savetxt(out_file, np.transpose(np.vstack([t, values])), delimiter=",")
        for i in range(self.NUM_DRONES):
            s = self.states[i]
            for stem, row in _DIRECT:
                put(stem + str(i), s[row, :])
            for stem, row in _RATES:
                put(stem + str(i), np.hstack([0, (s[row, 1:] - s[row, 0:-1]) * self.LOGGING_FREQ_HZ]))
            for m in range(4):
                put("rpm" + str(m) + "-" + str(i), s[12 + m, :])
                put("pwm" + str(m) + "-" + str(i), (s[12 + m, :] - 4070.3) / 0.2685)
        return csv_dir

    def plot(self, pwm=False):
        raise NotImplementedError("plotting needs matplotlib: out of scope; use save()/save_as_csv() and plot offline")

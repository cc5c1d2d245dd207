"""
Sensor constants of the simulated LiDARs.

API mirror (same class names, field names, defaults and factory names) of the reference's
``lidar/lidar_intrinsics.py``:

    LidarIntrinsics              :12-25
    DualAxisLidarIntrinsics      :28-211   (BLK2GO preset :152-186)
    Indoor8LineLidarIntrinsics   :214-389  (presets :246-350)

These are plain host-side parameter records; the rays they describe are generated on the GPU
(see ``sensors.py`` and ``csrc/scan.cu``).
"""
from __future__ import annotations

import math
from abc import ABC
from dataclasses import dataclass, field
from typing import List, Optional, Tuple

import numpy as np

_DEG = math.pi / 180.0
_TWO_PI = 2.0 * math.pi


@dataclass
class LidarIntrinsics(ABC):
    """Fields every sensor has (reference lidar_intrinsics.py:12-25).  fov_* are positive degrees."""

    fov_up: float
    fov_down: float
    vertical_res: int
    horizontal_res: int
    max_range: float
    vertical_degrees: List[float] = None


def _ladder(n_lines: int, top: float = 15.0, span: float = 35.0) -> List[float]:
    """n_lines elevations from ``top`` down to ``top - span`` degrees, rounded to 0.1 deg
    (reference lidar_intrinsics.py:273-276, :297-300)."""
    return [round(top - (i * span / (n_lines - 1)), 1) for i in range(n_lines)]


@dataclass
class Indoor8LineLidarIntrinsics(LidarIntrinsics):
    """Single-axis spinning multi-line sensor (reference lidar_intrinsics.py:214-243)."""

    fov_up: float = 15.0
    fov_down: float = 20.0
    vertical_res: int = 8
    horizontal_res: int = 2000
    max_range: float = 20.0
    vertical_degrees: List[float] = field(default_factory=lambda: [15, 10, 5, 0, -5, -10, -15, -20])

    min_range: float = 0.1
    range_resolution: float = 0.01
    scan_frequency: float = 10.0
    points_per_beam: int = 2000

    range_noise_std: float = 0.02
    angle_noise_std: float = 0.01

    dual_axis: bool = False
    capture_rate: int = 200000
    intensity_noise_std: float = 0.1
    dropout_probability: float = 0.05

    # ---- presets: (keyword overrides) per reference factory ---------------------------------------
    _PRESETS = {
        "standard_8line": {},                                                             # :246-248
        "high_resolution_8line": dict(horizontal_res=4000, points_per_beam=4000,          # :251-257
                                      range_resolution=0.005),
        "low_cost_8line": dict(horizontal_res=1000, points_per_beam=1000,                 # :260-267
                               range_resolution=0.02, range_noise_std=0.05),
    }

    @classmethod
    def _preset(cls, name: str) -> "Indoor8LineLidarIntrinsics":
        return cls(**cls._PRESETS[name])

    @classmethod
    def create_standard_8line(cls) -> "Indoor8LineLidarIntrinsics":
        return cls._preset("standard_8line")

    @classmethod
    def create_high_resolution_8line(cls) -> "Indoor8LineLidarIntrinsics":
        return cls._preset("high_resolution_8line")

    @classmethod
    def create_low_cost_8line(cls) -> "Indoor8LineLidarIntrinsics":
        return cls._preset("low_cost_8line")

    @classmethod
    def create_dense_32line(cls) -> "Indoor8LineLidarIntrinsics":
        """32 lines x 4000 azimuths, 25 m (reference :270-289) -- BASELINE config C2's sensor."""
        return cls(fov_up=15.0, fov_down=20.0, vertical_res=32, horizontal_res=4000, max_range=25.0,
                   vertical_degrees=_ladder(32), points_per_beam=3000, range_resolution=0.005,
                   range_noise_std=0.01, angle_noise_std=0.005)

    @classmethod
    def create_leica_blk2go(cls) -> "Indoor8LineLidarIntrinsics":
        """64-line single-axis stand-in for the BLK2GO (reference :292-317)."""
        return cls(fov_up=15.0, fov_down=20.0, vertical_res=64, horizontal_res=8000, max_range=25.0,
                   vertical_degrees=_ladder(64), points_per_beam=5000, range_resolution=0.003,
                   range_noise_std=0.003, angle_noise_std=0.002, min_range=0.5, scan_frequency=20.0,
                   dual_axis=True, capture_rate=420000)

    @classmethod
    def create_custom_lidar(cls, num_beams: int = 8, beam_angles: Optional[List[float]] = None,
                            horizontal_resolution: float = 0.1, max_range: float = 20.0,
                            points_per_beam: int = 2000) -> "Indoor8LineLidarIntrinsics":
        """User-defined beam table; azimuth count = 360/horizontal_resolution capped at 10000 (reference :320-350)."""
        if beam_angles:
            up, down, table = max(beam_angles), abs(min(beam_angles)), beam_angles
        else:
            up, down, table = 15.0, 20.0, [15, 10, 5, 0, -5, -10, -15, -20]
        width = min(int(360.0 / horizontal_resolution), 10000)
        return cls(fov_up=up, fov_down=down, vertical_res=num_beams, horizontal_res=width, max_range=max_range,
                   vertical_degrees=table, points_per_beam=points_per_beam)

    def get_total_points_per_scan(self) -> int:
        return self.vertical_res * self.horizontal_res          # reference :352-354

    def get_scan_frequency(self) -> float:
        return self.scan_frequency

    def get_range_limits(self) -> tuple:
        return (self.min_range, self.max_range)

    def add_noise(self, points: np.ndarray, ranges: np.ndarray, angles: np.ndarray, intensities: np.ndarray) -> tuple:
        """Post-hoc measurement noise on already-computed arrays (reference :364-389).  The reference never
        calls this from the simulator (it is off the hot path); kept for API completeness, host numpy."""
        rng = np.random
        ranges_n = ranges + rng.normal(0, self.range_noise_std, ranges.shape)
        angles_n = angles + rng.normal(0, np.deg2rad(self.angle_noise_std), angles.shape)
        inten_n = np.clip(intensities + rng.normal(0, self.intensity_noise_std, intensities.shape), 0, 1)
        if self.dropout_probability > 0:
            keep = rng.random(len(points)) > self.dropout_probability
            return points[keep], ranges_n[keep], angles_n[keep], inten_n[keep]
        return points, ranges_n, angles_n, inten_n


@dataclass
class DualAxisLidarIntrinsics(LidarIntrinsics):
    """Dual-axis (rotating + swinging) scanner such as the Leica BLK2GO (reference lidar_intrinsics.py:28-66)."""

    fov_up: float = 15.0
    fov_down: float = 20.0
    vertical_res: int = 1
    horizontal_res: int = 1
    max_range: float = 25.0
    vertical_degrees: List[float] = None

    phi_0: float = 0.0
    omega_phi: float = _TWO_PI

    scan_duration: float = 1.0
    point_rate: int = 420000

    phi_range: tuple = (0.0, _TWO_PI)
    theta_range: tuple = (-20.0 * np.pi / 180, 15.0 * np.pi / 180)

    angle_noise_std: float = 0.001
    timing_jitter_std: float = 0.0001
    dropout_probability: float = 0.02

    frame_duration: float = 0.1
    num_vertical_lines: int = 32

    swing_amplitude: float = 5.0 * np.pi / 180
    swing_frequency: float = 1.0

    def get_scan_parameters(self) -> dict:
        keys = ("phi_0", "omega_phi", "scan_duration", "point_rate", "phi_range", "theta_range",
                "swing_amplitude", "swing_frequency")
        return {k: getattr(self, k) for k in keys}

    def base_theta_angles(self) -> np.ndarray:
        """Rest elevation of each line, top to bottom (reference indoor_lidar.py:247-249)."""
        return np.linspace(self.theta_range[1], self.theta_range[0], self.num_vertical_lines)

    def calculate_angles_at_time(self, t: float, line_idx: int = 0) -> tuple:
        """(phi, theta) of one line at time t for the time-parameterised model (reference :80-118).
        Off the hot path (the simulator uses the azimuth-parameterised table of DualAxisLidar.get_rays)."""
        phi = (self.phi_0 + self.omega_phi * t) % _TWO_PI
        base = self.base_theta_angles()[line_idx % self.num_vertical_lines]
        phase = line_idx * _TWO_PI / self.num_vertical_lines
        theta = float(np.clip(base + self.swing_amplitude * np.sin(self.swing_frequency * t + phase),
                              self.theta_range[0], self.theta_range[1]))
        if self.angle_noise_std > 0:
            phi += np.random.normal(0, self.angle_noise_std)
            theta += np.random.normal(0, self.angle_noise_std)
        return phi, theta

    def generate_time_sequence(self, frame_duration: float = None) -> np.ndarray:
        """Uniform sample times of one output frame (reference :120-138)."""
        span = self.frame_duration if frame_duration is None else frame_duration
        n = int(self.point_rate * span)
        return np.arange(0, span, span / n)

    def get_total_points_per_scan(self) -> int:
        return int(self.point_rate * self.scan_duration)        # reference :140-142

    def get_scan_frequency(self) -> float:
        return 1.0 / self.scan_duration

    def get_range_limits(self) -> tuple:
        return (0.5, self.max_range)

    @classmethod
    def create_blk2go_dual_axis(cls) -> "DualAxisLidarIntrinsics":
        """640 kpt/s x 0.1 s = 64000 rays per frame on 32 swinging lines (reference :152-186).
        The simulator's default sensor (reference s3dis_simulator.py:59-60,605)."""
        return cls(fov_up=15.0, fov_down=20.0, vertical_res=1, horizontal_res=1, max_range=25.0, vertical_degrees=None,
                   phi_0=0.0, omega_phi=_TWO_PI, scan_duration=0.1, point_rate=640000,
                   phi_range=(0.0, _TWO_PI), theta_range=(-20.0 * np.pi / 180, 15.0 * np.pi / 180),
                   angle_noise_std=0.001, timing_jitter_std=0.0001, dropout_probability=0.02,
                   frame_duration=0.1, num_vertical_lines=32, swing_amplitude=5.0 * np.pi / 180, swing_frequency=1.0)

    @classmethod
    def create_custom_dual_axis(cls, phi_0: float = 0.0, theta_0: float = 15.0, omega_phi: float = _TWO_PI,
                                omega_theta: float = -0.1, point_rate: int = 420000,
                                scan_duration: float = 1.0) -> "DualAxisLidarIntrinsics":
        """Custom dual-axis sensor.  The reference's version (:188-211) forwards ``theta_0``, ``omega_theta``
        and ``use_spiral_scan`` to a dataclass that has no such fields and therefore always raises TypeError;
        here those three are accepted and ignored so the factory is usable."""
        del theta_0, omega_theta
        return cls(phi_0=phi_0, omega_phi=omega_phi, scan_duration=scan_duration, point_rate=point_rate,
                   frame_duration=0.1, fov_up=15.0, fov_down=20.0, vertical_res=1, horizontal_res=1, max_range=25.0)

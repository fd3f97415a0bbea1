"""Sensor-data containers (farms_core ``AnimatData`` / ``SensorsData`` stand-ins).

The reference logs into ``data.sensors.<kind>.array`` -- float64 arrays shaped
``[buffer_size, n_items, n_cols]`` with a ``names`` list
(farms_mujoco/simulation/task.py:98-101,158,208-216; sensors.pyx:149-156;
swimming/drag.pyx:189-191,265-267).  The batched engine keeps the same
per-environment layout and adds a leading environment axis:
``[n_envs, buffer_size, n_items, n_cols]``; ``data.env(i)`` is a view with the
reference's exact shape.
"""

import numpy as np

from .layout import sc


class SensorArray:
    """``names`` + ``array`` pair (``LinkSensorArray`` & co. stand-in)."""

    def __init__(self, names, array):
        self.names = list(names)
        self.array = array

    def size(self, dim):
        return self.array.shape[dim]


class SensorsData:
    """``data.sensors`` (links / joints / contacts / xfrc / muscles)."""

    def __init__(self, links, joints, contacts, xfrc, muscles=None):
        self.links = links
        self.joints = joints
        self.contacts = contacts
        self.xfrc = xfrc
        self.muscles = muscles if muscles is not None else SensorArray([], np.zeros((0, 0, 0)))


class AnimatData:
    """Per-environment log with the reference's array shapes."""

    def __init__(self, timestep, sensors):
        self.timestep = timestep
        self.sensors = sensors

    @classmethod
    def from_sensors_names(cls, timestep, buffer_size, links, joints, contacts=(), xfrc=(),
                           muscles=(), dtype=np.float64):
        """Mirror of ``AnimatData.from_sensors_names`` (task.py:208-216)."""
        assert not muscles, 'muscles are outside the hot path (SURVEY.md section 2 row 4)'

        def zeros(n, cols):
            return np.zeros((buffer_size, n, cols), dtype=dtype)

        sensors = SensorsData(
            links=SensorArray(links, zeros(len(links), sc.link_size)),
            joints=SensorArray(joints, zeros(len(joints), sc.joint_size)),
            contacts=SensorArray(contacts, zeros(len(contacts), sc.contact_size)),
            xfrc=SensorArray(xfrc, zeros(len(xfrc), sc.xfrc_size)),
        )
        return cls(timestep=timestep, sensors=sensors)


class BatchedAnimatData:
    """``[n_envs, buffer_size, n_items, n_cols]`` logs sharing one ``names`` set."""

    def __init__(self, timestep, n_envs, buffer_size, links, joints, contacts=(), xfrc=(),
                 dtype=np.float32):
        self.timestep = timestep
        self.n_envs = n_envs
        self.buffer_size = buffer_size

        def zeros(n, cols):
            return np.zeros((n_envs, buffer_size, n, cols), dtype=dtype)

        self.sensors = SensorsData(
            links=SensorArray(links, zeros(len(links), sc.link_size)),
            joints=SensorArray(joints, zeros(len(joints), sc.joint_size)),
            contacts=SensorArray(contacts, zeros(len(contacts), sc.contact_size)),
            xfrc=SensorArray(xfrc, zeros(len(xfrc), sc.xfrc_size)),
        )

    def env(self, index):
        """Reference-shaped view of one environment (no copy)."""
        s = self.sensors
        sensors = SensorsData(
            links=SensorArray(s.links.names, s.links.array[index]),
            joints=SensorArray(s.joints.names, s.joints.array[index]),
            contacts=SensorArray(s.contacts.names, s.contacts.array[index]),
            xfrc=SensorArray(s.xfrc.names, s.xfrc.array[index]),
        )
        return AnimatData(timestep=self.timestep, sensors=sensors)

"""Contact sensors (mirror of farms_mujoco/sensors/sensors.pyx).

``cycontacts2data`` (sensors.pyx:140-190) walks MuJoCo's active contacts, calls
``mj_contactForce`` and accumulates reaction / friction / total force and the
force-weighted position per contact sensor.  Here that loop runs inside the team kernel
(csrc/fb_device.h, FbStep::write_log) for every environment, and the per-thread kernel
writes the zero rows of contact-free steps; this function copies the finished row from
the device log.  Muscle sensors (sensors.pyx:193-297) need farms_muscle: out of scope.
"""


def cycontacts2data(physics, iteration, data, geompair2data=None, meters=1.0, newtons=1.0):
    """Contacts to data: ``data.array[:, iteration]`` <- device log row (all environments).

    ``geompair2data`` / ``meters`` / ``newtons`` are accepted for signature parity; they were
    compiled into the engine's tables (FbFarms.cand_sensor, units) at construction."""
    del geompair2data, meters, newtons
    if data.array.shape[2]:
        data.array[:, iteration] = physics.log_row('contacts', iteration, physics.log_stride > 1)

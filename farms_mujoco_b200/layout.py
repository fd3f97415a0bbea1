"""Sensor-array column layout (farms_core ``sensor_convention.sc`` stand-in).

farms_core is a third-party dependency that is absent from the reference tree
and from this image, so its column enum is restated here **in one place**.
What the reference itself pins (farms_mujoco/simulation/physics.py:423-524,
farms_mujoco/sensors/sensors.pyx:39-51, farms_mujoco/swimming/drag.pyx:189-191,
265-267): link columns 0..2 are the CoM position, the named link ranges are
each contiguous, contacts have 12 columns in reaction/friction/total/position
order, xfrc has 6 columns force|torque.  The numeric values of the remaining
joint columns are a recollection of farms_core (SURVEY.md Appendix C,
UNVERIFIED); ``check_against_farms_core`` compares them whenever farms_core is
importable.
"""


class _SensorConvention:
    """Column indices into ``data.sensors.<kind>.array[iteration, item, :]``."""

    # Links -- 20 columns
    link_size = 20
    link_com_position_x = 0
    link_com_position_y = 1
    link_com_position_z = 2
    link_com_orientation_x = 3
    link_com_orientation_y = 4
    link_com_orientation_z = 5
    link_com_orientation_w = 6
    link_urdf_position_x = 7
    link_urdf_position_y = 8
    link_urdf_position_z = 9
    link_urdf_orientation_x = 10
    link_urdf_orientation_y = 11
    link_urdf_orientation_z = 12
    link_urdf_orientation_w = 13
    link_com_velocity_lin_x = 14
    link_com_velocity_lin_y = 15
    link_com_velocity_lin_z = 16
    link_com_velocity_ang_x = 17
    link_com_velocity_ang_y = 18
    link_com_velocity_ang_z = 19

    # Joints -- 18 columns
    joint_size = 18
    joint_position = 0
    joint_velocity = 1
    joint_force_x = 2
    joint_force_y = 3
    joint_force_z = 4
    joint_torque_x = 5
    joint_torque_y = 6
    joint_torque_z = 7
    joint_cmd_position = 8
    joint_cmd_velocity = 9
    joint_cmd_torque = 10
    joint_torque = 11
    joint_torque_active = 12
    joint_torque_stiffness = 13
    joint_torque_damping = 14
    joint_torque_friction = 15
    joint_limit_force = 16
    joint_reserved = 17

    # Contacts -- 12 columns (sensors.pyx:39-51)
    contact_size = 12
    contact_reaction_x = 0
    contact_reaction_y = 1
    contact_reaction_z = 2
    contact_friction_x = 3
    contact_friction_y = 4
    contact_friction_z = 5
    contact_total_x = 6
    contact_total_y = 7
    contact_total_z = 8
    contact_position_x = 9
    contact_position_y = 10
    contact_position_z = 11

    # External forces -- 6 columns (drag.pyx:265-267)
    xfrc_size = 6
    xfrc_force_x = 0
    xfrc_force_y = 1
    xfrc_force_z = 2
    xfrc_torque_x = 3
    xfrc_torque_y = 4
    xfrc_torque_z = 5


sc = _SensorConvention()

# Upper-case aliases used by the Cython side of the reference (sensors.pyx:39-51)
CONTACT_REACTION_X, CONTACT_REACTION_Y, CONTACT_REACTION_Z = 0, 1, 2
CONTACT_FRICTION_X, CONTACT_FRICTION_Y, CONTACT_FRICTION_Z = 3, 4, 5
CONTACT_TOTAL_X, CONTACT_TOTAL_Y, CONTACT_TOTAL_Z = 6, 7, 8
CONTACT_POSITION_X, CONTACT_POSITION_Y, CONTACT_POSITION_Z = 9, 10, 11


def layout_words():
    """The four row widths as the C-ABI expects them (links, joints, contacts, xfrc)."""
    return sc.link_size, sc.joint_size, sc.contact_size, sc.xfrc_size


def check_against_farms_core():
    """Compare this table with farms_core when it is importable.

    Returns a list of mismatching attribute names ([] if equal), or None when
    farms_core is not installed (the situation in this image).
    """
    try:
        from farms_core.sensors.sensor_convention import sc as ref_sc  # noqa
    except Exception:  # pragma: no cover - farms_core absent in this image
        return None
    bad = []
    for name in dir(_SensorConvention):
        if name.startswith('_'):
            continue
        if hasattr(ref_sc, name) and int(getattr(ref_sc, name)) != int(getattr(sc, name)):
            bad.append(name)
    return bad

"""Drop-in for the reference's ``Driver_Models.py`` (reference ``Driver_Models.py:2-9``).

Named by the north star next to the VAE path, but it carries no arithmetic of that path and
nothing in the reference imports it: it stays plain Python with the reference's behaviour,
including the ZeroDivisionError when both vehicles move at the same speed.
"""

BRAKE_DECEL = 6.0        # m/s^2, the deceleration the rule commands (returned negated)
REACTION_MARGIN = 0.35   # s


def Reg157(x_ego, v_ego, x_front, v_front):
    """UN-R157-style braking rule: time-to-collision against a speed-dependent threshold.
    Returns -6 (brake) when ``ttc > v_rel / 12 + 0.35``, else None - the comparison and its
    direction are the reference's, kept as is."""
    closing_speed = v_ego - v_front
    time_to_collision = abs(x_front - x_ego) / closing_speed
    if time_to_collision > closing_speed / (2 * BRAKE_DECEL) + REACTION_MARGIN:
        return -6
    return None

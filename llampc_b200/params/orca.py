"""ORCA 1:43 car parameters -- same keys and values as llampc/params/orca.py:8-84 of the reference."""

_NOMINAL = (("lf", 0.029), ("lr", 0.033), ("mass", 0.041), ("Iz", 27.8e-6),
            ("Bf", 2.579), ("Br", 3.3852), ("Cf", 1.2), ("Cr", 1.2691), ("Df", 0.192), ("Dr", 0.1737),
            ("Cm1", 0.287), ("Cm2", 0.0545), ("Cr0", 0.0518), ("Cr2", 0.00035))
_LIMITS = (("max_acc", 5.0), ("min_acc", -5.0), ("max_pwm", 1.0), ("min_pwm", -0.1),
           ("max_steer", 0.35), ("min_steer", -0.35), ("max_steer_vel", 5.0))


def ORCA(control="pwm"):
    """Parameter dict for ``Dynamic(**params)``; ``control`` selects which input limits are listed
    ("pwm": duty cycle + steering, "acc": acceleration + steering), as in the reference."""
    p = dict(_NOMINAL)
    p.update(_LIMITS)
    if control == "pwm":
        hi, lo = p["max_pwm"], p["min_pwm"]
    elif control == "acc":
        hi, lo = p["max_acc"], p["min_acc"]
    else:
        raise NotImplementedError('choose control as "pwm" for Dynamic model and "acc" for Kinematic model')
    p["max_inputs"] = [hi, p["max_steer"]]
    p["min_inputs"] = [lo, p["min_steer"]]
    p["max_rates"] = [None, p["max_steer_vel"]]
    p["min_rates"] = [None, -p["max_steer_vel"]]
    return p

"""IDM vehicle record and the two parameter recipes (reference: road/vehicle/micro_vehicle.py:20-122)."""
import numpy as np

from road.vehicle.vehicle import DEFAULT_VEHICLE_LENGTH, Vehicle

IDM_FIELDS = ("accel_max", "accel_pref", "target_speed", "min_space", "time_pref", "length")   # rows of params[6][V]


class MicroVehicle(Vehicle):
    def __init__(self, id, position, speed, accel_max, accel_pref, target_speed, min_space, time_pref, length, a):
        super().__init__(id, position, speed, length, a)
        self.accel_max, self.accel_pref, self.target_speed = accel_max, accel_pref, target_speed
        self.min_space, self.time_pref = min_space, time_pref

    def idm_params(self):
        return [float(getattr(self, k)) for k in IDM_FIELDS]

    @staticmethod
    def default_micro_vehicle(speed_limit: float):
        """Parameters tied to the speed limit: a_max = u_max, a_pref = 0.8 u_max, v_target = 0.9 u_max,
        s0 = 0.1 len, T = 0.1 (micro_vehicle.py:30-72); id -1, parked at 0 until placed."""
        ln = DEFAULT_VEHICLE_LENGTH
        return MicroVehicle(-1, 0, 0, speed_limit * 1.0, speed_limit * 0.8, speed_limit * 0.9, ln * 0.1, 0.1, ln, ln)

    @staticmethod
    def random_micro_vehicle(speed_limit: float):
        """Uniform draws, in the reference's order of np.random calls (micro_vehicle.py:88-109)."""
        ln = DEFAULT_VEHICLE_LENGTH
        lo_hi = ((1.5 * speed_limit, 2.0 * speed_limit), (1.0 * speed_limit, 1.5 * speed_limit),
                 (0.8 * speed_limit, 1.2 * speed_limit), (0.2 * ln, 0.4 * ln), (0.2, 0.6))
        a_max, a_pref, v_t, s0, t_pref = (float(np.interp(np.random.rand(), [0, 1], list(b))) for b in lo_hi)
        return MicroVehicle(-1, 0, 0, a_max, a_pref, v_t, s0, t_pref, ln, ln)

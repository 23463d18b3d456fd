"""Vehicle attribute record (reference: road/vehicle/vehicle.py:1-17)."""
DEFAULT_VEHICLE_LENGTH = 5.0


class Vehicle:
    """id, position and speed along the lane, length, and `a`: the ancillary unit of "mass" that
    carries the density gradient of the cell a vehicle was spawned from (conversion.py:60-62)."""

    def __init__(self, id, position, speed, length, a):
        self.id, self.position, self.speed, self.length, self.a = id, position, speed, length, a

"""Lane shell shared by macro and micro lanes (reference: road/lane/_base_lane.py:7-54)."""
from typing import Dict


def synced(name):
    """Attribute stored under `name` whose every read / write first executes the network's queued steps."""
    def get(self):
        self._sync()
        return self.__dict__[name]

    def put(self, value):
        if "_net" in self.__dict__:
            self._sync()
        self.__dict__[name] = value
    return property(get, put)


class BaseLane:
    def __init__(self, id: int, length: float, speed_limit: float):
        self._net = None        # the RoadNetwork this lane was added to: owner of the queue of deferred steps
        self.id = id
        self.length = length
        self.speed_limit = speed_limit
        self.next_lane: Dict[int, "BaseLane"] = {}
        self.prev_lane: Dict[int, "BaseLane"] = {}

    def _sync(self):
        """Run the steps the network has queued (dropin/deferred.py) before lane state is read or changed."""
        net = self._net
        if net is not None and net._pending:
            net.flush()

    # kind / step hooks: filled in by MacroLane and MicroLane
    def is_macro(self):
        raise NotImplementedError()

    def is_micro(self):
        raise NotImplementedError()

    def forward(self, delta_time: float):
        raise NotImplementedError()

    def update_state(self):
        raise NotImplementedError()

    def clear(self):
        raise NotImplementedError()

    # adjacency
    def add_prev_lane(self, lane):
        self.prev_lane[lane.id] = lane

    def add_next_lane(self, lane):
        self.next_lane[lane.id] = lane

    def num_prev_lane(self):
        return len(self.prev_lane)

    def num_next_lane(self):
        return len(self.next_lane)

    def has_prev_lane(self):
        return bool(self.prev_lane)

    def has_next_lane(self):
        return bool(self.next_lane)

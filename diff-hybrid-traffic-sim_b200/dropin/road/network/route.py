"""Topology inputs of the network step (reference: road/network/route.py:3-82)."""
from typing import Dict, List


class MacroRoute:
    """Which lane currently feeds / drains each macro lane: two id->id maps, -1 when unconnected."""

    def __init__(self):
        self.next_lane_dict: Dict[int, int] = {}
        self.prev_lane_dict: Dict[int, int] = {}

    def get_next_lane(self, lane_id: int) -> int:
        return self.next_lane_dict.get(lane_id, -1)

    def get_prev_lane(self, lane_id: int) -> int:
        return self.prev_lane_dict.get(lane_id, -1)


class MicroRoute:
    """Lane-id sequence of one vehicle plus a cursor at the lane it is on."""

    def __init__(self, route: List[int], curr_idx: int = 0):
        self.route = route
        self.curr_idx = curr_idx

    def increment_curr_idx(self):
        self.curr_idx += 1

    def route_length(self) -> int:
        return len(self.route)

    def _at(self, k: int) -> int:
        return self.route[k] if 0 <= k < len(self.route) else -1

    def curr_lane_id(self) -> int:
        return self.route[self.curr_idx]

    def prev_lane_id(self) -> int:
        return self._at(self.curr_idx - 1)

    def next_lane_id(self) -> int:
        return self._at(self.curr_idx + 1)

"""
Constraints: the container the planner reads problem sizes, the goal-region
half-widths and the feasibility callback from (API of reference
lqrrt/constraints.py:31-61, same attribute names and error messages).

`is_feasible` must be a `lqrrt_b200.models.Feasibility` spec for planning to
run on the GPU; any callable is accepted here, as in the reference, and the
check happens when a Planner is built from it.
"""
import numpy as np


class Constraints(object):
    """
    nstates / ncontrols: dimensionality of state and effort.
    goal_buffer: half-edge lengths of the box defining the goal region.
    is_feasible: callable(x, u) -> bool (a models.Feasibility spec).
    """

    def __init__(self, nstates, ncontrols, goal_buffer, is_feasible):
        self.nstates = nstates
        self.ncontrols = ncontrols
        self.set_buffers(goal_buffer)
        self.set_feasibility_function(is_feasible)

    def set_buffers(self, goal_buffer=None):
        """Arguments not given are not modified (constraints.py:39-49)."""
        if goal_buffer is None:
            return
        if len(goal_buffer) != self.nstates:
            raise ValueError("The goal_buffer must have same dimensionality as state.")
        self.goal_buffer = np.abs(goal_buffer).astype(np.float64)

    def set_feasibility_function(self, is_feasible):
        """constraints.py:53-61."""
        if not hasattr(is_feasible, '__call__'):
            raise ValueError("Expected is_feasible to be a function.")
        self.is_feasible = is_feasible

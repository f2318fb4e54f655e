"""
lqrrt_b200 -- B200-native tree-expansion engine behind the lqRRT planner API.

Drop-in for the hot path of jnez71/lqRRT: `Planner`, `Constraints` and `Tree`
keep the reference's names, arguments and result attributes
(reference lqrrt/__init__.py:1-2 exports Constraints and Planner); the expand
loop runs as CUDA kernels (csrc/) reached through the C ABI in
include/lqrrt_b200.h.  `models` holds the device model specs that take the
place of the reference's Python callbacks (`plugin` compiles user-supplied
dynamics into the engine at run time), `configs` the demo and benchmark
problems.
"""
from .constraints import Constraints
from .planner import Planner
from .tree import Tree
from . import models, configs, waves, plugin

__all__ = ['Constraints', 'Planner', 'Tree', 'models', 'configs', 'waves', 'plugin']
__version__ = '0.1.0'

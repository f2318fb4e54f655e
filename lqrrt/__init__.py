"""
`import lqrrt` shim: the reference's package name (lqrrt/__init__.py:1-2)
resolved to the B200 engine, so code written against jnez71/lqRRT only has to
swap its model callbacks for `lqrrt.models` specs.
"""
from lqrrt_b200 import Constraints, Planner, Tree, models, configs, plugin  # noqa: F401

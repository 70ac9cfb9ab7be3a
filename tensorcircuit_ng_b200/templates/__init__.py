"""The slice of `tensorcircuit.templates` that sits directly on the statevector path."""
from . import graphs, measurements  # noqa: F401

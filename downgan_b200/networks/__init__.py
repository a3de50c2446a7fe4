from .generator import Generator  # noqa: F401
from .critic import Critic  # noqa: F401

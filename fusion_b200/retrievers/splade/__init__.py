from .base import BaseModel  # noqa: F401
from .splade import SPLADE  # noqa: F401

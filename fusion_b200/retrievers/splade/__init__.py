from .base import BaseModel  # noqa: F401

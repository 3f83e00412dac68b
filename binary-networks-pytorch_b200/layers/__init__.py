from .binary_modules import Conv1d, Conv2d, Linear, NotLowerable

__all__ = ["Linear", "Conv2d", "Conv1d"]

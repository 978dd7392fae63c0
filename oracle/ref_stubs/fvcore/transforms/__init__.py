from .transform import *  # noqa

from .module import *  # noqa: F401,F403

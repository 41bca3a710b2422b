from .loss_fn import *  # noqa: F401,F403

"""Import shim (test infrastructure only): cupyx.scipy.ndimage -> scipy.ndimage."""
from scipy.ndimage import *  # noqa: F401,F403

"""Version of the Snekmer API this package mirrors (snekmer/_version.py:1) and of the B200 library."""
__version__ = "1.3.0"
__b200_version__ = "0.1.0"

"""Stand-in for the un-vendored third-party package `transforms3d` (SpinRelax requirements.txt:5,
unpinned). Only used to import the reference in the build container (oracle/ref_loader.py)."""

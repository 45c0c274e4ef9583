"""Empty stand-in so that `import mdtraj as md` at calculate-Ct-from-traj.py:5 succeeds when only the
pure-numpy functions of that script are needed (oracle/ref_loader.py)."""

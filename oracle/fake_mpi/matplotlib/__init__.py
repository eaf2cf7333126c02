"""Empty stand-in for matplotlib (pcg_solver.py:17 imports pyplot, never used when PlotFlag is False)."""

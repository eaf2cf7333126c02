"""Empty stand-in for matplotlib.pyplot (oracle harness only)."""

"""Empty stand-in: only needed because reference examples import the plotting Logger."""

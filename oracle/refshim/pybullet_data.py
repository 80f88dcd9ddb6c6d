"""Stand-in for `pybullet_data` (reference BaseAviary.py:13,482). Test infrastructure only."""
def getDataPath():
    return "."

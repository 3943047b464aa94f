"""Stand-in for ``scikit-image`` (requirements.txt:12, absent offline): ``skimage.metrics.structural_similarity`` only."""

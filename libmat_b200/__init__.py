"""libmat_b200: B200-native RPD3D + dist2mat hot path behind LibMAT's host entry points."""

"""TEST INFRASTRUCTURE ONLY -- import stand-in: sam/datasets/_image_features_reader.py imports h5py at module level;
nothing on the SA-M4C hot path reads feature files."""

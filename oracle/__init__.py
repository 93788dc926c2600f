"""CPU oracle = TEST INFRASTRUCTURE (see oracle_np.py header).  Never imported by the product package."""

from vision4leg_b200.algo.rl_algo import RLAlgo   # noqa: F401

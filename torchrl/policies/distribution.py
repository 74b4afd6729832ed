from vision4leg_b200.policies.distribution import TanhNormal   # noqa: F401

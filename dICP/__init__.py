"""Drop-in alias: `from dICP.ICP import ICP`, `from dICP.nn import nn`, `from dICP.loss import loss`
resolve to the B200-native implementation in `dicp_b200` (same names, same signatures)."""

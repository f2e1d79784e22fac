from dicp_b200.ICP import ICP          # noqa: F401
from dicp_b200.nn import nn            # noqa: F401
from dicp_b200.loss import loss        # noqa: F401

from dicp_b200.nn import nn            # noqa: F401

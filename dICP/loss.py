from dicp_b200.loss import loss        # noqa: F401

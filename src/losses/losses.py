from nextgen_uia_b200.losses import InfoNCELoss  # noqa: F401

"""Drop-in `src` package: the reference's import paths (`from src.adapters import ...`,
`from src.losses import InfoNCELoss`) resolved onto the B200 implementation in nextgen_uia_b200."""

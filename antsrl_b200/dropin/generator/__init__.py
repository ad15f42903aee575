"""Drop-in mirror of the reference's ``generator`` package."""

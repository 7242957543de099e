"""Stand-in for matplotlib: inert mocks (plots are not part of the parity surface)."""

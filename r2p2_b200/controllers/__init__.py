"""Controller plugins with a device form (same module / class names as r2p2/controllers/)."""

"""Device versions of the reference demos named by BASELINE.json `configs`."""

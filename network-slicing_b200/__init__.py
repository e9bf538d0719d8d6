"""ranslice-b200: batched RAN-slicing env step path for B200."""

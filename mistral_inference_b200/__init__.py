"""B200-native drop-in for the mistral-inference transformer hot path."""

"""B200-native DeepFM embedding + interaction hot path."""

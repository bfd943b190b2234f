"""B200-native nano_gicp scan-registration hot path."""

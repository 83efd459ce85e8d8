"""gradient-boosted-normalizing-flows_b200: B200-native boosted-mixture density path.  Import as `gbnf_b200`."""

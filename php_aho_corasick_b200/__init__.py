"""B200-native Aho-Corasick matcher behind the C boundary of ph4r05/php_aho_corasick."""
__all__ = ["native", "workloads"]

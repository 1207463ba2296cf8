"""sopht_mpi_b200: B200-native hot path of sopht-mpi behind the reference operator API."""

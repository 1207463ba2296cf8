"""Rank-filtered logger with the method set of the reference's ``MPILogger``
(reference ``sopht_mpi/utils/mpi_logger.py:63-158``) but without MPI: the rank is
taken from the process-group environment (``RANK``) when present."""
import logging
import os
import sys


class RankLogger:
    def __init__(self, echo_rank=(0,), level=logging.WARNING):
        self.rank = int(os.environ.get("RANK", "0"))
        self.echo_rank = list(echo_rank)
        self._logger = logging.getLogger(f"rank[{self.rank}]")
        if not self._logger.handlers:
            handler = logging.StreamHandler(sys.stdout)
            handler.setFormatter(
                logging.Formatter("%(asctime)s %(name)s %(levelname)s: %(message)s")
            )
            self._logger.addHandler(handler)
        self._logger.setLevel(level)
        self._logger.propagate = False

    def set_log_level(self, level):
        self._logger.setLevel(level)

    def set_echo_rank(self, echo_rank):
        self.echo_rank = list(echo_rank)

    def _emit(self, fn, msg):
        if self.rank in self.echo_rank:
            fn(msg)

    def debug(self, msg):
        self._emit(self._logger.debug, msg)

    def info(self, msg):
        self._emit(self._logger.info, msg)

    def warning(self, msg):
        self._emit(self._logger.warning, msg)

    def error(self, msg):
        self._emit(self._logger.error, msg)

    def critical(self, msg):
        self._emit(self._logger.critical, msg)


logger = RankLogger(echo_rank=[0])

import logging


def _log_api_usage(identifier):
    pass


def setup_logger(output=None, distributed_rank=0, *, name="detectron2", **kw):
    logger = logging.getLogger(name)
    return logger


def create_small_table(d):
    return str(d)


def log_every_n_seconds(lvl, msg, n=1, *, name=None):
    logging.getLogger(name or __name__).log(lvl, msg)

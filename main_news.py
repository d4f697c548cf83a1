"""Adressa entry point: the flags and period lists of the reference's main_news.py (63 periods, online training from
period 21, testing from 48; MF_epochs=2 TR_epochs=2 multi_num=7) on the B200 path."""
from main_yelp import get_parse as _get_parse, main as _main


def get_parse():
    return _get_parse("news")


if __name__ == "__main__":
    _main("news")

"""``utils.read_balfile`` (reference: utils/read_balfile.py:4-37): same 9-tuple."""
from gbp_b200.balio import read_bal


def read_balfile(balfile):
    return read_bal(balfile).as_tuple()

"""Reference-named package ``gbp`` (see gbp_b200.run): same import surface as joeaortiz/gbp."""
from . import gbp
from . import gbp_ba

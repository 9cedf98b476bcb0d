from . import lie_algebra
from . import transformations
from . import derivatives
from . import read_balfile
from . import gaussian

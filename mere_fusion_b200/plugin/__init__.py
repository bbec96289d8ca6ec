"""Host-side mirror of the reference's plugin surface (BaseReal / BaseASR subclasses) for the three
heads: same class names, constructor and method signatures, queue and (chunk, type) tuple
contracts as /root/reference/{basereal,baseasr,lipreal,lipasr,musereal,museasr,nerfreal,nerfasr}.py, with the model
call replaced by the sm_100a engines behind the C ABI and the per-session child process replaced
by an in-process thread."""
from .baseasr import BaseASR            # noqa: F401
from .basereal import BaseReal          # noqa: F401
from .lipasr import LipASR              # noqa: F401
from .lipreal import LipReal            # noqa: F401
from .museasr import MuseASR            # noqa: F401
from .musereal import MuseReal          # noqa: F401
from .nerfasr import NerfASR            # noqa: F401
from .nerfreal import NeRFReal          # noqa: F401

"""Drop-in modules named after the reference's: put this directory on sys.path and the reference's
`from utils import *`, `from utils_lr import *`, `from my_losses import *` resolve here."""

"""UNetFamily — drop-in package name of the reference (jcfszxc/jcfszxc-UNet) so that
`from UNetFamily import UNet; UNet.UNet()` and pickled `UNetFamily.UNet.UNet` objects resolve
(reference train.py:28-44,374,502).  The modules execute on the B200-native kernels."""

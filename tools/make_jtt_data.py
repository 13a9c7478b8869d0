"""Development tool (needs /root/reference): exports the JTT amino-acid rate matrix and frequencies the reference
embeds (``/root/reference/pastml/models/JTTModel.py:31-61``, from Kosiol & Goldman's Dayhoff-matrix page) into
``pastml_b200/models/data/jtt.npz`` in the reference's state order (one-letter codes, sorted)."""
import os
import sys

import numpy as np

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
from oracle import ref_env  # noqa: E402

ref_env.install()
from pastml.models.JTTModel import JTT_STATES, JTT_FREQUENCIES, JTT_RATE_MATRIX  # noqa: E402

out = os.path.join(REPO, 'pastml_b200', 'models', 'data', 'jtt.npz')
np.savez(out, states=np.array(JTT_STATES), frequencies=np.array(JTT_FREQUENCIES, dtype=np.float64),
         rate_matrix=np.array(JTT_RATE_MATRIX, dtype=np.float64))
print(out, JTT_STATES)

"""Host-side mirrors of the reference's global modules needed at the hot-path boundary
(reference: src/lisflood/global_modules/)."""

from tapqir_b200.imscroll.glimpse_reader import GlimpseDataset, bin_hist, read_glimpse  # noqa: F401

"""stub: the reference imports matplotlib at module level; nothing is plotted."""

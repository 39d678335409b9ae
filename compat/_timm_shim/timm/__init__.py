"""Stand-in for the three timm symbols the reference imports (timm.models.layers.{DropPath, to_2tuple, trunc_normal_}),
for containers without timm.  Not on the product path."""

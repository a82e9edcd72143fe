/* empty stand-in for the missing external/smesh submodule */

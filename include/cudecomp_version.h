/* Version macros of the C ABI this library implements.
 * Mirrors reference include/cudecomp_version.h:20-27 (library 0.7.0, all struct layouts at version 1). */
#ifndef CUDECOMP_VERSION_H
#define CUDECOMP_VERSION_H

#define CUDECOMP_MAJOR 0
#define CUDECOMP_MINOR 7
#define CUDECOMP_PATCH 0

#define CUDECOMP_GRID_DESC_CONFIG_VERSION 1
#define CUDECOMP_GRID_DESC_AUTOTUNE_OPTIONS_VERSION 1
#define CUDECOMP_PENCIL_INFO_VERSION 1

#endif

/* host/octrn.h -- plain-C reader (and writer, for tests and stand-in assets) of the `.octrn` asset files the reference
 * reads through its binary-only octarine libraries (octarine_mesh_read_from_file / octarine_image_read_from_file,
 * main.c:526-559; on-disk layout: SURVEY.md App. B, restated in malevich_b200/assets.py).
 *
 *   all files : "eniratco" (8 bytes), u32 asset type (0 image, 1 mesh), u32 reserved
 *   mesh      : u32 size_of_data, u32 vertex_count, u32 index_count, then 32-byte vertices (pos3, normal3, uv2) and
 *               u32 indices -- exactly what load_mesh (main.c:526-536) points p_vertex_buffer / p_index_buffer at
 *   image     : u64 size_of_data, u32 format, u16 width, height, depth, array_size, mip_levels, flags, then texels
 *               (format 0x4801 = R32G32B32A32_FLOAT, anything else = 4 bytes per texel)
 */
#ifndef MALEVICH_OCTRN_H
#define MALEVICH_OCTRN_H
#include <stddef.h>
#include <stdint.h>

#define OCTRN_FORMAT_R32G32B32A32_FLOAT 0x4801u
#define OCTRN_FORMAT_R8G8B8A8_UNORM 0x1c01u /* what this writer stores for 8-bit textures; the reader accepts any non-float format */

typedef struct OctrnMeshHeader { /* MeshHeader main.c:54-58 */
	uint32_t size, vertex_count, index_count;
} OctrnMeshHeader;

typedef struct OctrnImageHeader {
	uint64_t size;
	uint32_t format;
	uint16_t width, height, depth, array_size, mip_levels, flags;
} OctrnImageHeader;

/* 0 on success; *pp_data is malloc'ed (vertices followed by indices / texels) and owned by the caller. */
int octrn_read_mesh(const char *path, OctrnMeshHeader *header, void **pp_data);
int octrn_read_image(const char *path, OctrnImageHeader *header, void **pp_data);
int octrn_write_mesh(const char *path, const void *vertices, uint32_t vertex_count, const uint32_t *indices, uint32_t index_count);
int octrn_write_image(const char *path, uint32_t format, uint16_t width, uint16_t height, const void *texels);
const char *octrn_last_error(void);

#endif

/* host/octrn.c -- see octrn.h */
#include "octrn.h"

#include <stdio.h>
#include <stdlib.h>
#include <string.h>

static char g_error[256] = "";
const char *octrn_last_error(void) { return g_error; }

static int fail(const char *path, const char *what) {
	snprintf(g_error, sizeof(g_error), "%s: %s", path, what);
	return 1;
}

/* whole file into memory; the common 16-byte header is checked here */
static unsigned char *slurp(const char *path, uint32_t want_type, size_t *out_size) {
	FILE *f = fopen(path, "rb");
	if(!f) {
		fail(path, "cannot open");
		return NULL;
	}
	fseek(f, 0, SEEK_END);
	const long size = ftell(f);
	fseek(f, 0, SEEK_SET);
	unsigned char *blob = size >= 16 ? (unsigned char *)malloc((size_t)size) : NULL;
	if(!blob || fread(blob, 1, (size_t)size, f) != (size_t)size) {
		fclose(f);
		free(blob);
		fail(path, "short file");
		return NULL;
	}
	fclose(f);
	uint32_t type;
	memcpy(&type, blob + 8, 4);
	if(memcmp(blob, "eniratco", 8) != 0 || type != want_type) {
		free(blob);
		fail(path, "not an octarine asset of the expected type");
		return NULL;
	}
	*out_size = (size_t)size;
	return blob;
}

int octrn_read_mesh(const char *path, OctrnMeshHeader *header, void **pp_data) {
	size_t size;
	unsigned char *blob = slurp(path, 1u, &size);
	if(!blob) return 1;
	if(size < 28) {
		free(blob);
		return fail(path, "mesh header truncated");
	}
	memcpy(header, blob + 16, 12);
	const size_t payload = (size_t)header->vertex_count * 32 + (size_t)header->index_count * 4;
	if(payload != header->size || 28 + payload != size) {
		free(blob);
		return fail(path, "mesh header inconsistent with the file size");
	}
	memmove(blob, blob + 28, payload); /* the caller gets the payload at the start of the allocation */
	*pp_data = blob;
	return 0;
}

int octrn_read_image(const char *path, OctrnImageHeader *header, void **pp_data) {
	size_t size;
	unsigned char *blob = slurp(path, 0u, &size);
	if(!blob) return 1;
	if(size < 40) {
		free(blob);
		return fail(path, "image header truncated");
	}
	memcpy(&header->size, blob + 16, 8);
	memcpy(&header->format, blob + 24, 4);
	memcpy(&header->width, blob + 28, 12);
	const size_t texel = header->format == OCTRN_FORMAT_R32G32B32A32_FLOAT ? 16 : 4;
	if(40 + header->size != size || (size_t)header->width * header->height * texel > header->size) {
		free(blob);
		return fail(path, "image header inconsistent with the file size");
	}
	memmove(blob, blob + 40, (size_t)header->size);
	*pp_data = blob;
	return 0;
}

static int write_head(FILE *f, uint32_t type) {
	const uint32_t zero = 0;
	return fwrite("eniratco", 1, 8, f) == 8 && fwrite(&type, 4, 1, f) == 1 && fwrite(&zero, 4, 1, f) == 1;
}

int octrn_write_mesh(const char *path, const void *vertices, uint32_t vertex_count, const uint32_t *indices, uint32_t index_count) {
	FILE *f = fopen(path, "wb");
	if(!f) return fail(path, "cannot create");
	const OctrnMeshHeader h = { vertex_count * 32u + index_count * 4u, vertex_count, index_count };
	const int ok = write_head(f, 1u) && fwrite(&h, 12, 1, f) == 1 && fwrite(vertices, 32, vertex_count, f) == vertex_count && fwrite(indices, 4, index_count, f) == index_count;
	fclose(f);
	return ok ? 0 : fail(path, "short write");
}

int octrn_write_image(const char *path, uint32_t format, uint16_t width, uint16_t height, const void *texels) {
	FILE *f = fopen(path, "wb");
	if(!f) return fail(path, "cannot create");
	const size_t texel = format == OCTRN_FORMAT_R32G32B32A32_FLOAT ? 16 : 4;
	const uint64_t size = (uint64_t)width * height * texel;
	const uint16_t rest[6] = { width, height, 1, 1, 1, 0 };
	const int ok = write_head(f, 0u) && fwrite(&size, 8, 1, f) == 1 && fwrite(&format, 4, 1, f) == 1 && fwrite(rest, 2, 6, f) == 6 && fwrite(texels, 1, (size_t)size, f) == (size_t)size;
	fclose(f);
	return ok ? 0 : fail(path, "short write");
}

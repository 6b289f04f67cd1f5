/*
 * d3h_mesh.h -- C ABI of the triangle-mesh stage that consumes the extraction.
 *
 * D3-Human wraps every extracted surface in `render/mesh.py:Mesh` three to five times per iteration
 * (geometry/hmsdf.py:460-484, 554-593).  Each constructor runs
 *     Mesh.get_edge       render/mesh.py:240-250   sort the 3*F face edges, torch.unique(dim=0)  -> (E,2) int64
 * and each mesh then goes through
 *     auto_normals        render/mesh.py:418-446   area-weighted face normals scatter-added to the vertices,
 *                                                  degenerate -> (0,0,1), normalised; differentiable w.r.t. v_pos
 * The reference has no native code for either (plain PyTorch ops); the entry points below replace them behind the
 * same Python names (d3human-code_b200/render/mesh.py).  Conventions are those of d3h_tets.h: raw device pointers,
 * 0 / negative D3H_E_* return codes with d3h_last_error_string(), caller-owned memory, no host synchronisation,
 * everything enqueued on `stream`.
 */
#ifndef D3H_MESH_H_
#define D3H_MESH_H_

#include <stdint.h>

#include "d3h_tets.h"

#ifdef __cplusplus
extern "C" {
#endif

/* Result sizes of d3h_mesh_edges (device copy + optional pinned, device-mapped host copy; `seq` is written last). */
typedef struct d3h_mesh_counts {
  int64_t n_edges;   /* E: distinct undirected edges */
  int64_t bad_index; /* != 0: some face index was outside [0, n_verts); those edges were skipped */
  int64_t overflow;  /* != 0: E > cap_edges, rows beyond the capacity were dropped */
  int64_t seq;       /* echo of the caller's tag */
} d3h_mesh_counts;

/* Scratch of d3h_mesh_edges for a mesh of n_faces triangles over n_verts vertices (256-byte aligned regions). */
int64_t d3h_mesh_edges_workspace_bytes(int64_t n_faces, int64_t n_verts);

/* Mesh.get_edge (render/mesh.py:240-250): the distinct undirected edges (min, max) of `faces`, in the ascending
 * lexicographic order torch.unique(dim=0) returns, as int64 rows.
 *   faces        (n_faces,3) int64, device
 *   edges        (cap_edges,2) int64, device, 16-byte aligned; E <= 3*n_faces always fits cap_edges = 3*n_faces
 *   counts_dev   device copy of the sizes (required)
 *   counts_host  pinned + device-mapped host copy polled with d3h_mesh_wait_counts, or NULL
 * E is published as soon as the edges are counted, before they are sorted and written. */
int d3h_mesh_edges(const int64_t* faces, int64_t n_faces, int64_t n_verts, int64_t* edges, int64_t cap_edges,
                   void* workspace, int64_t workspace_bytes, d3h_mesh_counts* counts_dev,
                   d3h_mesh_counts* counts_host, int64_t seq, d3h_stream_t stream);

/* Spin until counts_host->seq == seq (the GPU writes it through mapped memory); D3H_E_TIMEOUT after timeout_us. */
int d3h_mesh_wait_counts(const d3h_mesh_counts* counts_host, int64_t seq, int64_t timeout_us);

/* auto_normals forward (render/mesh.py:418-441).
 *   pos     (n_verts,3) fp32         faces (n_faces,3) int64
 *   v_nrm   (n_verts,3) fp32 out     acc   (n_verts,4) fp32 out, 16-byte aligned: the un-normalised sums (x,y,z,0),
 *                                          kept by the caller as the tape of the backward call
 *   bad     optional device int32: set to 1 if a face index is outside [0, n_verts) (such faces are skipped)
 * With exactly three faces the reference's torch.cross (no `dim`) crosses along the face axis; reproduced. */
int d3h_mesh_normals_forward(const float* pos, const int64_t* faces, int64_t n_verts, int64_t n_faces, float* v_nrm,
                             float* acc, int32_t* bad, d3h_stream_t stream);

/* auto_normals backward: g_pos (n_verts,3) is overwritten with d loss / d pos given g_nrm (n_verts,3). */
int d3h_mesh_normals_backward(const float* pos, const int64_t* faces, int64_t n_verts, int64_t n_faces,
                              const float* acc, const float* g_nrm, float* g_pos, d3h_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* D3H_MESH_H_ */

"""Host-side mirror of the part of the reference's `render/` package that wraps the extracted surfaces
(SURVEY.md section 8(f) row 2): `render.mesh.Mesh` and `render.mesh.auto_normals`."""

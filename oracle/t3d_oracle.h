/*
 * t3d_oracle.h -- CPU oracle for the Tetra3D software geometry stage.
 *
 * TEST INFRASTRUCTURE ONLY.  This is a plain-C restatement of the reference's
 * Go algorithm (SolarLune/tetra3d, Camera.Render -> Model.ProcessVertices ->
 * light.Light -> emit).  Only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs may load it.  The product (libt3dgeo.so,
 * libt3dhost.so) never links, loads or calls anything in oracle/.
 *
 * PARITY STATUS: "parity unpinned" for the hot path itself.  The reference holds
 * no golden vectors for ProcessVertices / lighting / sorting;
 * the only assertion-bearing test in the reference, TestMatrixInversion
 * (matrix_test.go:36-55), is reproduced as a known-answer test against
 * orc_mat4_* (tests/test_oracle_kat.py).  No Go toolchain exists in the build
 * image, so the reference cannot be executed here.  A second restatement of the
 * path, written separately and shaped like the Go source (tests/literal_go.py),
 * agrees with this one bit for bit on the parity scenes and on scenes built from
 * the reference's own GLB assets (tests/test_literal_restatement.py).  Everything else follows the
 * cited reference lines operation by operation: float32 everywhere, no FMA
 * contraction (-ffp-contract=off), left-to-right evaluation as Go/amd64
 * (GOAMD64=v1) does.
 *
 * All file:line citations are relative to the reference tree.
 */
#ifndef T3D_ORACLE_H
#define T3D_ORACLE_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---- enums (material.go:9-55, world.go:7-11) ------------------------------ */
enum { ORC_SORT_BACK_TO_FRONT = 0, ORC_SORT_FRONT_TO_BACK = 1, ORC_SORT_NONE = 2 };
enum { ORC_TRANSPARENCY_AUTO = 0, ORC_TRANSPARENCY_OPAQUE = 1, ORC_TRANSPARENCY_ALPHA_CLIP = 2,
       ORC_TRANSPARENCY_TRANSPARENT = 3 };
enum { ORC_BILLBOARD_Y_UP_GLOBAL_Y = 0, ORC_BILLBOARD_Y_UP_CAMERA_Y = 1 };
enum { ORC_LIGHTING_DEFAULT = 0, ORC_LIGHTING_FIXED_NORMALS = 1, ORC_LIGHTING_DOUBLE_SIDED = 2 };
enum { ORC_DEPTH_MODE_DEFAULT = 0, ORC_DEPTH_MODE_UNBILLBOARDED = 1 };
enum { ORC_LV_PER_VERTEX_WITHOUT_NORMAL = 0, ORC_LV_PER_VERTEX_WITH_NORMAL = 1, ORC_LV_PER_OBJECT = 2,
       ORC_LV_OFF = 3 };
enum { ORC_FOG_ADD = 0, ORC_FOG_SUB = 1, ORC_FOG_OVERWRITE = 2 };
enum { ORC_LIGHT_AMBIENT = 0, ORC_LIGHT_POINT = 1, ORC_LIGHT_DIRECTIONAL = 2, ORC_LIGHT_VOLUME = 3 };

/* ---- scene description (flat, borrowed pointers) --------------------------- */

/* Node local TRS (node.go:406, 572-608).  rotation is a row-major Matrix4. */
typedef struct {
    float   position[3];
    float   scale[3];
    float   rotation[16];
    int32_t parent;              /* index into orc_scene.nodes, -1 = none */
} orc_node;

/* Material fields the geometry stage reads (material.go:66-131). */
typedef struct {
    int32_t present;             /* MeshPart.Material != nil */
    float   color[4];
    int32_t backface_culling;
    int32_t triangle_sort_mode;
    int32_t shadeless;
    int32_t lighting_mode;
    int32_t light_volume_shading_mode;
    int32_t billboard_enabled, billboard_lock_x, billboard_lock_y, billboard_lock_z;
    int32_t billboard_up_direction;
    int32_t billboarded_depth_mode;
    int32_t transparency_mode;
    int32_t blend_source_over;   /* mat.Blend == ebiten.BlendSourceOver */
    int32_t visible;
    int32_t render_order;
    int32_t use_texture;         /* UseTexture && Texture != nil */
    float   tex_w, tex_h;        /* Texture.Bounds().Dx/Dy */
} orc_material;

/* MeshPart (mesh.go:2183-2190); triangle_end is INCLUSIVE as in the reference. */
typedef struct {
    int32_t vertex_index_start, vertex_index_end;
    int32_t triangle_start, triangle_end;
    orc_material material;
} orc_part;

/* Mesh SoA arrays (mesh.go:260-302). */
typedef struct {
    int32_t n_verts;
    const float*    positions;      /* 3*n */
    const float*    normals;        /* 3*n */
    const float*    uvs;            /* 2*n */
    const float*    colors;         /* 4*n: the ACTIVE channel, or NULL */
    int32_t         active_color_channel; /* Mesh.VertexActiveColorChannel (<0: none) */
    const float*    weights;        /* 4*n or NULL (mesh.VertexWeights, padded to 4) */
    const uint16_t* bones;          /* 4*n or NULL (mesh.VertexBones, padded to 4)   */
    const uint8_t*  n_influences;   /* n or NULL: len(VertexWeights[v])              */
    int32_t n_tris;
    const uint32_t* indices;        /* 3*t: Triangle.VertexIndexA/B/C */
    int32_t n_parts;
    const orc_part* parts;
    int32_t n_shape_keys;           /* mesh.go:796-822 */
    const float* key_positions;     /* k*n*3 */
    const float* key_normals;       /* k*n*3 */
    const float* key_weights;       /* k */
    /* Auto-subdivision (mesh.go:447-485, 1758-1959).  When auto_subdivide is set, `indices` hold every base triangle
     * followed by its pre-built sub-triangles, level by level: the order MeshPart.forEachTri(true, ...) visits them in
     * (mesh.go:2238-2256), and the arrays below describe every entry. */
    int32_t auto_subdivide;
    const float*   tri_centers;     /* 3*t Triangle.Center (sub-triangles: mesh.go:1828-1830), overrides RecalculateCenter */
    const int32_t* tri_sub_base;    /* t: the base triangle's index (a base triangle names itself) */
    const int8_t*  tri_sub_level;   /* t: -1 = base triangle, else its level in the base's subdivisionLevels */
    const uint8_t* tri_n_levels;    /* t: len(subdivisionLevels) of a base triangle */
    const float*   tri_max_span;    /* t: Triangle.MaxSpan */
    const float*   tri_normal;      /* 3*t: Triangle.Normal (base triangles) */
} orc_mesh;

/* Model (model.go:20-61). */
typedef struct {
    int32_t node;
    int32_t mesh;
    float   color[4];
    int32_t shadeless;
    int32_t visible;
    int32_t frustum_culling;
    int32_t skinned;
    int32_t n_bones;
    const float* bone_influences;   /* n_bones*16: Node.boneInfluence (node.go:592-594) per skin joint */
    int32_t light_group;            /* index into light_groups, -1 = nil */
    int32_t batch_owner;            /* Model.DynamicBatchOwner: model index, -1 = nil */
    int32_t batch_owner_part;       /* the owner's MeshPart key in DynamicBatchModels */
} orc_model;

/* ILight (light.go:12-30): ambient/point/directional/volume. */
typedef struct {
    int32_t type;
    int32_t node;
    float   color[3];
    float   energy;
    int32_t visible;
    float   range;                  /* PointLight.lightRange */
    int32_t volume;                 /* index into volumes for ORC_LIGHT_VOLUME */
} orc_light;

typedef struct {
    int32_t n_lights;
    const int32_t* lights;          /* indices into orc_scene.lights */
    int32_t active;
} orc_light_group;

/* LightVolume lookup data (light.go:376-396, 583-663, 717-728, 1267-1272). */
typedef struct {
    float dim_min[3], dim_max[3];
    float cell_size[3];
    int32_t cell_count[3];
    float offset[3];
    float light_bias[3];
    const float* cell_colors;       /* 3 * cx*cy*cz, index x + y*cx + z*cx*cy */
    /* LightVolumeCell.useLights (light.go:1381, 1508-1514): per cell (same index as cell_colors) a run of scene light
     * indices, in SetUseLights order; cell_light_start == NULL: no cell uses lights */
    const int32_t* cell_light_start;    /* [cells + 1] */
    const int32_t* cell_light_index;    /* indices into orc_scene.lights */
} orc_light_volume;

/* Camera (camera.go:259-332). */
typedef struct {
    int32_t node;
    int32_t width, height;
    float   field_of_view;
    float   near_plane, far_plane;
    int32_t perspective;
    float   ortho_scale;
    float   depth_margin;
    float   vertex_snapping;
    int32_t render_depth;
    int32_t render_normals;
    int32_t perspective_corrected_texture_mapping;
    int32_t max_light_count;        /* Camera.MaxLightCount (camera.go:1474); 0 = unlimited */
    int32_t max_light_focus_node;   /* Camera.MaxLightCountFocusPoint as a node index, or -1 = the camera (camera.go:1476-1481) */
} orc_camera;

/* World (world.go:25-42). */
typedef struct {
    int32_t present;
    int32_t lighting_on;
    int32_t fog_on;
    int32_t fog_mode;
    float   fog_color[4];
    float   fog_range[2];
    float   ambient_color[3];
    float   ambient_energy;
    int32_t ambient_visible;
} orc_world;

typedef struct {
    int32_t n_nodes;        const orc_node* nodes;
    int32_t n_meshes;       const orc_mesh* meshes;
    int32_t n_models;       const orc_model* models;        /* DFS order of RenderNodes (camera.go:1200-1210) */
    int32_t n_lights;       const orc_light* lights;        /* scene lights, DFS order */
    int32_t n_scene_lights; const int32_t* scene_lights;    /* the subset that is in the tree under root */
    int32_t n_light_groups; const orc_light_group* light_groups;
    int32_t n_volumes;      const orc_light_volume* volumes;
    int32_t n_subdivision_levels; const float* subdivision_levels;   /* scene.library.autosubdivisionLevels: {DistanceSquared, MaximumTriangleSize} pairs */
    orc_camera camera;
    orc_world  world;
} orc_scene;

/* ---- results ---------------------------------------------------------------- */

/* ebiten.Vertex (github.com/hajimehoshi/ebiten/v2 v2.9.7, go.mod:6): 12 x float32. */
typedef struct {
    float DstX, DstY, SrcX, SrcY, ColorR, ColorG, ColorB, ColorA, Custom0, Custom1, Custom2, Custom3;
} orc_vertex;

/* One render(rp) call (camera.go:1507-1827). */
typedef struct {
    int32_t model, part;
    int32_t flush;                 /* which flush() consumed it */
    int64_t first_vertex;          /* into the concatenated lists */
    int32_t vertex_count;          /* 3 * drawn triangles */
    int32_t total_tris;            /* meshPart.TriangleCount() */
    float   min_depth, max_depth;  /* model.go:710-711, 962-967 */
} orc_render_rec;

/* One flush(rp) call (camera.go:1829-2048). */
typedef struct {
    int32_t model, part;
    int64_t first_vertex;
    int32_t vertex_count;
} orc_flush_rec;

typedef struct {
    int32_t n_renders;  orc_render_rec* renders;
    int32_t n_flushes;  orc_flush_rec*  flushes;
    int64_t n_vertices;
    orc_vertex* color_list;
    orc_vertex* depth_list;
    orc_vertex* normal_list;
    uint16_t*   index_list;
    uint32_t*   tri_ids;           /* n_vertices/3: mesh triangle index of each emitted triangle */
    /* optional capture of the per-vertex intermediates of every render() call, concatenated in render
     * order over the part's vertex range [vertex_index_start, vertex_index_end): */
    int64_t n_captured;
    float*  cap_clip;              /* 4 per vertex: globalVertexTransforms  */
    float*  cap_light;             /* 4 per vertex: mesh.vertexLights.colors */
    /* DebugInfo counters (camera.go:51-57) */
    int32_t total_parts, drawn_parts, batched_parts, total_tris, drawn_tris, light_count, active_light_count;
    /* wall-clock split, seconds (camera.go:41-48 mirrors) */
    double t_total, t_process, t_light, t_emit;
} orc_result;

enum { ORC_OK = 0, ORC_E_UNSUPPORTED = -3, ORC_E_INVALID = -1 };

/* Runs Camera.RenderScene on the scene.  capture != 0 fills cap_clip/cap_light. */
int  orc_render_scene(const orc_scene* scene, int capture, orc_result** out);
void orc_free_result(orc_result* r);
/* Model.BakeLighting (model.go:1203-1268) of model `model` with the lights light_ids[0..n) plus the world's ambient light:
 * rgb = VertexColors[targetChannel] of its mesh, [n_verts][3], in / out. */
/* Model.BakeAO, same-model pass (model.go:1040-1124); Triangle.Normal / MaxSpan (mesh.go:2018-2037) are inputs. */
int  orc_bake_ao(const orc_mesh* mesh, const float* tri_normals3, const float* tri_max_span, const uint8_t* target_parts,
                 const uint8_t* target_vertices, float occlusion_angle, const float* occlusion_color4, float* colors4);
double orc_go_acos(double x);                                                             /* Go math.Acos */
int  orc_bake_lighting(const orc_scene* scene, int32_t model, const int32_t* light_ids, int32_t n, float* rgb);
const char* orc_last_error(void);

/* ---- math exposed for known-answer tests (matrix.go) ------------------------- */
void  orc_mat4_identity(float* m);
void  orc_mat4_mult(const float* a, const float* b, float* out);        /* matrix_amd64.s / matrix_none.go */
void  orc_mat4_inverted(const float* m, float* out);                     /* matrix.go:255-307 */
void  orc_mat4_transposed(const float* m, float* out);                   /* matrix.go:218-230 */
void  orc_mat4_rotate(float x, float y, float z, float angle, float* out); /* matrix.go:100-128 */
void  orc_mat4_translate(float x, float y, float z, float* out);
void  orc_mat4_scale(float x, float y, float z, float* out);
int   orc_mat4_equals(const float* a, const float* b);                   /* matrix.go:511-522 */
void  orc_mat4_decompose(const float* m, float* pos3, float* scale3, float* rot16); /* matrix.go:181-196 */
void  orc_mat4_lookat(const float* from3, const float* to3, const float* up3, float* out); /* matrix.go:788-815 */
void  orc_mat4_perspective(float fovy, float near_plane, float far_plane, float w, float h, float* out); /* matrix.go:588-604 */
void  orc_mat4_orthographic(float near_plane, float far_plane, float r, float l, float t, float b, float* out);
void  orc_node_transform(const orc_node* nodes, int32_t idx, float* out16); /* node.go:572-608 */
/* Triangle.RecalculateCenter (mesh.go:1972-1982) for every triangle of a mesh. */
void  orc_triangle_centers(const float* positions, const uint32_t* indices, int32_t n_tris, float* centers3);
/* The 512-bin assignment alone (sortingTriangles.go:38-64). */
int   orc_bin_index(float depth, float min_depth, float max_depth, int sort_mode);

/* ---- bone poses: the step before the path --------------------------------------------
 * AnimationTrack.ValueAsVector / ValueAsQuaternion (animation.go:63-164), Quaternion.Slerp / ToMatrix4
 * (quaternion.go:26-59, 132-178), AnimationPlayer.forceUpdate without relative motion (animation.go:837-980; the blend of
 * animation.go:852-907 is orc_pose_bones_blended), Node.Transform + boneInfluence (node.go:572-594).  Slerp goes through Go's float64
 * math.Sin / math.Atan2 (math32/math32.go:137-139, 204-206), which are pure Go on amd64 (Cephes-derived,
 * src/math/sin.go, atan.go, atan2.go of the Go standard library): restated below as orc_go_sin / orc_go_atan2. */
typedef struct orc_track { int32_t first_key, n_keys, interpolation; } orc_track;    /* 0 linear, 1 constant, 2 cubic (evaluated as linear: animation.go:107) */
typedef struct orc_skeleton {
    int32_t n_bones;
    const int32_t* parent;          /* [n] parent bone or -1 (child of the pose's root transform); parents come first */
    const float* inverse_bind;      /* [n][16] Node.inverseBindMatrix */
    const float* rest_position;     /* [n][3]  node.position / scale / rotation of bones the animation has no track for */
    const float* rest_scale;        /* [n][3] */
    const float* rest_rotation;     /* [n][16] */
} orc_skeleton;
typedef struct orc_animation {
    int32_t n_bones;
    const orc_track* tracks;        /* [n][3] position, scale, rotation */
    const float* key_times;
    const float* key_values;        /* [keys][4] */
} orc_animation;
/* boneInfluence of every bone, posed at `playhead`, with the skeleton's root bones parented to root16. */
void   orc_pose_bones(const orc_skeleton* sk, const orc_animation* an, float playhead, const float* root16, float* out);
/* The same while the player blends from `from` @ from_playhead into `an` @ playhead (animation.go:852-907), bp = blend fraction. */
void   orc_pose_bones_blended(const orc_skeleton* sk, const orc_animation* from, float from_playhead, const orc_animation* an, float playhead,
                              float bp, const float* root16, float* out);
/* Blends interrupted by further Play() calls: levels[0 .. n) in the order they were played, the last one playing; playheads[i] /
 * bps[i] are level i's values at its last update (bps[0] unused).  See pose_local_chain in t3d_oracle.c for what the reference's
 * currentProperties / prevAnimatedProperties maps make of such a history. */
void   orc_pose_bones_chain(const orc_skeleton* sk, int n_levels, const orc_animation* const* levels, const float* playheads, const float* bps,
                            const float* root16, float* out);
void   orc_pose_locals_chain(const orc_skeleton* sk, int n_levels, const orc_animation* const* levels, const float* playheads, const float* bps,
                             float* pos3, float* scale3, float* rot16, int32_t* set3);
void   orc_sample_bone(const orc_animation* an, int bone, float time, int32_t* flags3, float* pos3, float* scale3, float* quat4);
void   orc_quat_slerp(const float* a4, const float* b4, float percent, float* out4);      /* quaternion.go:26-59 */
void   orc_quat_to_matrix(const float* q4, float* out16);                                 /* quaternion.go:132-178 */
double orc_go_sin(double x);                                                              /* Go math.Sin (amd64: pure Go) */
double orc_go_atan2(double y, double x);                                                  /* Go math.Atan2 */

#ifdef __cplusplus
}
#endif
/* Test hook: perm = stable argsort of keys, computed with the restated sort.SliceStable (go_stable_sort). */
void orc_test_go_stable_sort(int n, const float* keys, int* perm);
/* The same under an arbitrary relation: less(i, j) = table[perm[i] * n + perm[j]] (not necessarily an ordering). */
void orc_test_go_stable_sort_table(int n, const uint8_t* table, int* perm);

#endif

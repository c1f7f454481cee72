/*
 * t3dgeo.h -- C ABI of libt3dgeo, the B200-native geometry stage for Tetra3D.
 *
 * The reference (SolarLune/tetra3d, pure Go) has no plugin / FFI interface for this path: the whole
 * geometry stage is the body of Model.ProcessVertices (model.go:600-995) plus the two
 * globalSortingTriangleBucket.ForEach loops of the `render` closure in Camera.Render
 * (camera.go:1610-1644 lighting, camera.go:1646-1820 emit), communicating through package globals
 * (tetra3d.go:21-35).  This header is the seam a cgo shim binds: Go keeps the
 * scene graph, the render-pair list and `flush` (camera.go:1829-2048); everything between "render
 * pairs are known" (camera.go:1457) and "flush consumes colorVertexList[:n]" runs on the GPU.
 *
 * Conventions
 *  - Plain C: pointers + sizes only.  Inputs are borrowed for the duration of the call (copied into
 *    library-owned pinned staging), so Go may move/free them afterwards (cgo pointer rules hold).
 *  - Outputs are library-owned buffers, valid until the next t3d_frame_begin on the same context.
 *  - Every function returns T3D_OK (0) or a negative T3D_E_* code; t3d_last_error gives the text.
 *    Nothing aborts.  Features that need host callbacks (Model.VertexTransformFunction model.go:47,
 *    Model.VertexClipFunction model.go:53, Material.CustomDepthFunction material.go:129) cannot run on the device: the shim
 *    must not submit such draws (it keeps the Go path for them) -- there is no CPU fallback in here.
 *  - One context per GPU, single caller at a time (the reference path is single-goroutine, non
 *    re-entrant: tetra3d.go:21-35, sortingTriangles.go:164).
 *  - All matrices are Tetra3D Matrix4: row-major float32[16], row-vector convention, translation in
 *    row 3 (matrix.go:9-10, 619-627).
 */
#ifndef T3DGEO_H
#define T3DGEO_H

#include <stdint.h>
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#define T3D_ABI_VERSION 6      /* 2: t3d_mesh_desc (auto-subdivision arrays) and t3d_volume_desc (cell lights) grew; 3: t3d_set_exact_lighting and the round-2 entry points; 4: t3d_submit_instances; 5: t3d_camera carries the frustum terms, t3d_submit_instances the culling bounds; 6: t3d_pose.from_playhead / blend (were reserved), t3d_submit_poses_blended, t3d_submit_pose_chains, t3d_set_async_download, t3d_frame_download_wait, t3d_set_skip_unused_depth_list */

enum {
    T3D_OK            =  0,
    T3D_E_INVALID     = -1,   /* bad argument / out-of-range index */
    T3D_E_CUDA        = -2,   /* CUDA runtime error (text in t3d_last_error) */
    T3D_E_UNSUPPORTED = -3,   /* feature needs a host callback (see header comment) */
    T3D_E_NOMEM       = -4,
    T3D_E_STATE       = -5,   /* call order violated (e.g. submit outside begin/execute) */
    T3D_E_NOTFOUND    = -6    /* unknown mesh / volume id */
};

/* Material.TriangleSortMode (material.go:9-13) */
enum { T3D_SORT_BACK_TO_FRONT = 0, T3D_SORT_FRONT_TO_BACK = 1, T3D_SORT_NONE = 2 };
/* Material.LightingMode (material.go:34-38) */
enum { T3D_LIGHTING_DEFAULT = 0, T3D_LIGHTING_FIXED_NORMALS = 1, T3D_LIGHTING_DOUBLE_SIDED = 2 };
/* Material.LightVolumeShadingMode (material.go:50-55) */
enum { T3D_LV_PER_VERTEX_WITHOUT_NORMAL = 0, T3D_LV_PER_VERTEX_WITH_NORMAL = 1, T3D_LV_PER_OBJECT = 2, T3D_LV_OFF = 3 };
/* World.FogMode (world.go:7-11) */
enum { T3D_FOG_ADD = 0, T3D_FOG_SUB = 1, T3D_FOG_OVERWRITE = 2 };
/* ILight implementations (light.go:80, 137, 290, 377) */
enum { T3D_LIGHT_AMBIENT = 0, T3D_LIGHT_POINT = 1, T3D_LIGHT_DIRECTIONAL = 2, T3D_LIGHT_VOLUME = 3 };

/* t3d_camera.flags */
enum {
    T3D_CAM_PERSPECTIVE      = 1u << 0,  /* camera.perspective */
    T3D_CAM_RENDER_DEPTH     = 1u << 1,  /* Camera.RenderDepth (camera.go:1744) */
    T3D_CAM_RENDER_NORMALS   = 1u << 2,  /* Camera.RenderNormals (camera.go:1710) */
    T3D_CAM_PERSP_CORRECT_UV = 1u << 3   /* Camera.PerspectiveCorrectedTextureMapping (camera.go:1689) */
};

/* t3d_draw.flags */
enum {
    T3D_DRAW_HAS_MATERIAL        = 1u << 0,   /* meshPart.Material != nil */
    T3D_DRAW_BACKFACE_CULLING    = 1u << 1,   /* Material.BackfaceCulling (model.go:902) */
    T3D_DRAW_LIT                 = 1u << 2,   /* `lighting` of camera.go:1521-1528 */
    T3D_DRAW_SKINNED             = 1u << 3,   /* model.skinned (model.go:715) */
    T3D_DRAW_BILLBOARD           = 1u << 4,   /* Material.BillboardEnabled (model.go:634) */
    T3D_DRAW_BILLBOARD_LOCK_X    = 1u << 5,
    T3D_DRAW_BILLBOARD_LOCK_Y    = 1u << 6,
    T3D_DRAW_BILLBOARD_LOCK_Z    = 1u << 7,
    T3D_DRAW_BILLBOARD_CAMERA_UP = 1u << 8,   /* BillboardUpDirection == BillboardYUpCameraY (model.go:639) */
    T3D_DRAW_DEPTH_UNBILLBOARDED = 1u << 9,   /* BillboardedDepthMode == DepthModeUnbillboarded (model.go:700) */
    T3D_DRAW_CONTINUE_FLUSH      = 1u << 10   /* dynamic batch child after the first: appends to the same
                                                 flush as the previous draw (camera.go:2082-2100) */
};

/* ebiten.Vertex (github.com/hajimehoshi/ebiten/v2 v2.9.7, go.mod:6): what colorVertexList /
 * depthVertexList / normalVertexList (tetra3d.go:21-23) hold.  Layout-identical so Go can
 * unsafe.Slice the returned pointer straight into DrawTriangles. */
typedef struct t3d_vertex {
    float DstX, DstY, SrcX, SrcY;
    float ColorR, ColorG, ColorB, ColorA;
    float Custom0, Custom1, Custom2, Custom3;
} t3d_vertex;

/* MeshPart (mesh.go:2183-2190).  tri_end is EXCLUSIVE here (Go's TriangleEnd is inclusive: pass
 * TriangleEnd+1).  The vertex range is informational; the library derives the range the part's
 * triangles actually reference. */
typedef struct t3d_part_range {
    uint32_t vtx_start, vtx_end;
    uint32_t tri_start, tri_end;
} t3d_part_range;

/* Mesh SoA arrays (mesh.go:260-302), uploaded once and whenever Go bumps `generation`. */
typedef struct t3d_mesh_desc {
    uint32_t        n_verts;
    const float*    positions;      /* 3*n  Mesh.VertexPositions                                     */
    const float*    normals;        /* 3*n  Mesh.VertexNormals                                       */
    const float*    uvs;            /* 2*n  Mesh.VertexUVs                                           */
    const float*    colors;         /* 4*n  Mesh.VertexColors[VertexActiveColorChannel] or NULL      */
    const float*    weights;        /* 4*n  Mesh.VertexWeights padded to 4 per vertex, or NULL       */
    const uint16_t* bones;          /* 4*n  Mesh.VertexBones padded to 4 per vertex, or NULL         */
    const uint8_t*  n_influences;   /* n    len(VertexWeights[v]) (<=4, gltf.go:675-700), or NULL=4  */
    uint32_t        n_tris;
    const uint32_t* indices;        /* 3*t  Triangle.VertexIndexA/B/C (mesh.go:1697-1699)            */
    const float*    centers;        /* 3*t  Triangle.Center (mesh.go:1980-1982)                      */
    uint32_t        n_parts;
    const t3d_part_range* parts;
    float           max_span;       /* Mesh.Dimensions.MaxSpan() (mesh.go:41-43; camera.go:1592)     */
    uint32_t        n_shape_keys;   /* mesh.shapeKeys (mesh.go:796-822)                              */
    const float*    key_positions;  /* k*n*3 */
    const float*    key_normals;    /* k*n*3 */
    /* Auto-subdivision (Mesh.SetAutoSubdivide mesh.go:447-485, performSubdivision mesh.go:1758-1861), optional (all NULL:
     * none).  `indices` / `centers` then hold every base triangle followed by its pre-built sub-triangles level by level,
     * the order MeshPart.forEachTri(true, ...) visits them in (mesh.go:2238-2256); the new vertices sit in the vertex
     * arrays like any other.  Which of {base, one level} is drawn is decided per draw on the device
     * (Triangle.handleSubdivision, mesh.go:1877-1959) from the levels of t3d_frame_set_autosubdivision. */
    const uint32_t* tri_sub_base;   /* t  index of the entry's base triangle (a base triangle names itself)           */
    const int8_t*   tri_sub_level;  /* t  -1 base triangle, else the level (index into subdivisionLevels)             */
    const uint8_t*  tri_n_levels;   /* t  len(subdivisionLevels) of a base triangle, 0 = never subdivided             */
    const float*    tri_max_span;   /* t  Triangle.MaxSpan (mesh.go:2018-2022)                                        */
    const float*    tri_normal;     /* 3*t Triangle.Normal (mesh.go:2030-2037), read for base triangles               */
} t3d_mesh_desc;

/* Per-frame camera uniforms (camera.go:1329-1341, 1471-1472, 1603). */
typedef struct t3d_camera {
    float    vp[16];          /* camera.ViewMatrix().Mult(camera.Projection())  (camera.go:1329)   */
    float    view[16];        /* camera.ViewMatrix() (model.go:707, RenderNormals only)            */
    float    position[3];     /* camera.WorldPosition() (model.go:624)                             */
    float    up[3];           /* camera.cameraUp (camera.go:1088; model.go:640)                    */
    float    near_plane, far_plane;   /* camera.near / camera.far (model.go:713-714)               */
    int32_t  width, height;   /* resultColorTexture bounds (camera.go:1471-1472)                   */
    float    depth_margin;    /* Camera.DepthMargin (camera.go:1339)                               */
    float    vertex_snapping; /* Camera.VertexSnapping (model.go:716, 807-812)                     */
    uint32_t flags;           /* T3D_CAM_*                                                          */
    /* What Camera.SphereInFrustum reads (camera.go:933-983), verbatim from the Camera's cached fields; only needed when
     * t3d_submit_instances is given culling bounds (zero otherwise).                                    */
    float    forward[3];      /* camera.cameraForward (camera.go:1086)                              */
    float    right[3];        /* camera.cameraRight (camera.go:1087); `up` above is cameraUp        */
    float    sphere_factor_y, sphere_factor_x, sphere_factor_tang;   /* camera.go:659-665           */
    float    ortho_scale;     /* camera.orthoScale (camera.go:967)                                  */
} t3d_camera;

/* scene.World fields the emit loop reads (camera.go:1784-1812). */
typedef struct t3d_world {
    int32_t present;          /* scene.World != nil */
    int32_t fog_on;
    int32_t fog_mode;
    float   fog_color[3];
    float   fog_range[2];
} t3d_world;

/* One entry of a light list as seen by one draw.  World-space data only: the per-model working
 * position / rotation of beginModel (light.go:175-189, 326-330) is derived on the device. */
typedef struct t3d_light {
    uint32_t type;            /* T3D_LIGHT_* */
    float    color[3];        /* LightBase.color RGB (light.go:34) */
    float    energy;          /* LightBase.energy */
    float    position[3];     /* light.WorldPosition() (point; light.go:184) */
    float    range;           /* PointLight.lightRange (light.go:141) */
    float    forward[3];      /* sun.WorldRotation().Forward() (light.go:323) */
    uint32_t volume;          /* handle from t3d_volume_upload (light volume) */
} t3d_light;

/* LightVolume lookup data (light.go:376-396, 583-663, 717-728, 1267-1272). */
typedef struct t3d_volume_desc {
    float    dim_min[3], dim_max[3];   /* l.dimensions */
    float    cell_size[3];
    uint32_t cell_count[3];
    float    offset[3];
    float    light_bias[3];
    const float* cell_colors;          /* 3 * cx*cy*cz, index x + y*cx + z*cx*cy (light.go:1271) */
    /* LightVolumeCell.useLights (light.go:1381, 1508-1514), optional: per cell a run of indices into the light array the
     * frame passes to t3d_volume_use_lights, in SetUseLights order.  NULL: no cell uses lights. */
    const uint32_t* cell_light_start;  /* [cells + 1] */
    const uint32_t* cell_light_index;  /* [cell_light_start[cells]] */
} t3d_volume_desc;

/* One render(rp) call (camera.go:1507): Model x MeshPart, in render order. */
typedef struct t3d_draw {
    uint64_t mesh_id;
    uint32_t part;
    uint32_t flags;                 /* T3D_DRAW_* */
    float    transform[16];         /* model.Transform() (model.go:616) */
    float    world_rotation[16];    /* model.WorldRotation() (light.go:328); used when lit by a sun */
    float    color[4];              /* mpColor = model.Color x Material.Color (camera.go:1597-1601) */
    float    src_w, src_h;          /* texture size or 0 (camera.go:1560-1566) */
    uint8_t  sort_mode;             /* T3D_SORT_* (camera.go:1540-1543; BackToFront when no material) */
    uint8_t  lighting_mode;         /* T3D_LIGHTING_* */
    uint8_t  light_volume_mode;     /* T3D_LV_* */
    uint8_t  reserved0;
    uint32_t bone_offset;           /* first matrix of this model in the t3d_submit_bones arena */
    uint32_t bone_count;
    uint32_t light_list;            /* handle returned by t3d_submit_lights */
    uint32_t reserved1;
} t3d_draw;

/* One of a run of Models that differ only in where they are and in their colour (particles and other DynamicBatch children,
 * instanced scenery): the Model's own Node.position / Node.scale (node.go:19-21) and Model.Color.  40 bytes here; what crosses
 * PCIe is 32 bytes per instance when every scale of the frame is uniform (x == y == z), 64 otherwise. */
typedef struct t3d_instance {
    float position[3];              /* node.position (local) */
    float scale[3];                 /* node.scale (local) */
    float color[4];                 /* model.Color */
} t3d_instance;

/* Per-draw result; `first_vertex` indexes the concatenated lists. */
typedef struct t3d_draw_result {
    uint64_t first_vertex;
    uint32_t vertex_count;          /* 3 * drawn triangles: what flush passes as list[:n] */
    uint32_t total_tris;            /* meshPart.TriangleCount() (camera.go:1532) */
    float    min_depth, max_depth;  /* model.go:710-711, 962-967 (+-MaxFloat32 when nothing survived) */
} t3d_draw_result;

/* CUDA-event timings of the last executed frame, milliseconds (DebugInfo mirrors, camera.go:41-48).
 * total_ms is always measured; the per-stage fields only when t3d_set_stage_timings(ctx, 1) was in force
 * for the frame (zero otherwise). */
typedef struct t3d_timings {
    float upload_ms;    /* always ~0: the uniforms are fetched over PCIe by the first kernel (part of prep_ms) */
    float prep_ms;      /* K0 fetch of the frame's uniforms + per-draw uniforms */
    float vertex_ms;    /* K1 fused skin/transform/light */
    float cull_ms;      /* K2 cull + depth + min/max */
    float sort_ms;      /* key + radix sort */
    float emit_ms;      /* K4 emit */
    float download_ms;  /* D2H of results (0 when results stayed on device) */
    float gather_ms;    /* t3d_gather_attach: end of the emit .. end of the delivery (shards: statistics + post; the issuing
                         * rank: waiting for and rebuilding every shard's lists); part of total_ms */
    float total_ms;     /* first kernel to last kernel (the device time of the frame) */
} t3d_timings;

typedef struct t3d_frame_output {
    uint64_t            n_vertices;     /* total emitted vertices over all draws */
    uint32_t            n_draws;
    const t3d_vertex*   color_list;     /* colorVertexList  (tetra3d.go:21) */
    const t3d_vertex*   depth_list;     /* depthVertexList  (tetra3d.go:23) */
    const t3d_vertex*   normal_list;    /* normalVertexList (tetra3d.go:22); NULL unless RENDER_NORMALS */
    const uint16_t*     index_list;     /* indexList (tetra3d.go:24): identity, uint16 wrap (camera.go:1816);
                                           a static table -- slice it [0, flush vertex count) */
    const uint32_t*     tri_ids;        /* n_vertices/3: mesh triangle index of each emitted triangle */
    const t3d_draw_result* draws;       /* n_draws, submit order */
} t3d_frame_output;

/* ---- bone poses evaluated on the device (the step before the path: SURVEY.md 8f N2) ---------------
 * A skeleton is the joint list of a skin (gltf.go:1452-1511): per bone its parent inside the list, its
 * Node.inverseBindMatrix and the local position / scale / rotation the bone's Node holds when no animation
 * track overrides them.  An animation holds, per bone, the three AnimationTracks of the bone's
 * AnimationChannel (animation.go:168-173): keyframe times and values.  With both resident, a frame ships one
 * t3d_pose (playhead + the transform the root bones hang from) per armature instead of n_bones matrices. */
enum { T3D_INTERPOLATION_LINEAR = 0, T3D_INTERPOLATION_CONSTANT = 1, T3D_INTERPOLATION_CUBIC = 2 /* evaluated as linear, animation.go:107 */ };
typedef struct t3d_skeleton_desc {
    uint32_t        n_bones;
    const int32_t*  parent;         /* [n] parent bone, or -1: child of t3d_pose.root; a parent precedes its children */
    const float*    inverse_bind;   /* [n][16] */
    const float*    rest_position;  /* [n][3] */
    const float*    rest_scale;     /* [n][3] */
    const float*    rest_rotation;  /* [n][16] */
} t3d_skeleton_desc;
typedef struct t3d_track { uint32_t first_key, n_keys, interpolation; } t3d_track;   /* n_keys == 0: the bone has no such track */
typedef struct t3d_animation_desc {
    uint32_t         n_bones;       /* bones [0, n_bones) of the skeleton it is played on */
    const t3d_track* tracks;        /* [n_bones][3]: TrackTypePosition, TrackTypeScale, TrackTypeRotation */
    uint32_t         n_keys;        /* keyframes over all tracks */
    const float*     key_times;     /* [n_keys] Keyframe.Time, non-decreasing inside a track */
    const float*     key_values;    /* [n_keys][4] Keyframe.Data: x y z _ or quaternion x y z w */
} t3d_animation_desc;
typedef struct t3d_pose {
    float root[16];                 /* Transform() of the node the skeleton's root bones are children of */
    float playhead;                 /* AnimationPlayer playhead the tracks are sampled at (animation.go:686-712) */
    float from_playhead;            /* t3d_submit_poses_blended: the playhead the previous animation was last sampled at */
    float blend;                    /* t3d_submit_poses_blended: `bp`, elapsed time / BlendTime clamped to 1 (animation.go:854-857) */
    float reserved;
} t3d_pose;

/* A player whose blends were interrupted by further Play() calls (t3d_submit_pose_chains): level i's animation sampled at
 * playhead[i]; blend[i] (i >= 1) the fraction of the blend into level i.  For the levels below the last, both are the values of the
 * level's last update before the next Play(). */
#define T3D_POSE_CHAIN_MAX 4
typedef struct t3d_pose_chain {
    float root[16];
    float playhead[T3D_POSE_CHAIN_MAX];
    float blend[T3D_POSE_CHAIN_MAX];    /* blend[0] is not read */
} t3d_pose_chain;

typedef struct t3d_ctx t3d_ctx;

/* ---- lifetime ---------------------------------------------------------------------------------- */
int         t3d_abi_version(void);
int         t3d_create(t3d_ctx** out, int device_id);
void        t3d_destroy(t3d_ctx* ctx);
const char* t3d_last_error(t3d_ctx* ctx);           /* ctx may be NULL (creation errors) */

/* ---- residency ---------------------------------------------------------------------------------- */
/* Replaces: the implicit "mesh lives in Go slices" of mesh.go:260-302.  Re-upload when dirty. */
int t3d_mesh_upload(t3d_ctx* ctx, uint64_t mesh_id, uint32_t generation, const t3d_mesh_desc* desc);
/* Updates only MeshShapeKey.weight values (mesh.go:796-822); k must equal n_shape_keys. */
int t3d_mesh_set_shape_key_weights(t3d_ctx* ctx, uint64_t mesh_id, const float* weights, uint32_t k);
int t3d_mesh_release(t3d_ctx* ctx, uint64_t mesh_id);
/* LightVolume grids (light.go:376-396). */
int t3d_volume_upload(t3d_ctx* ctx, uint32_t volume_id, const t3d_volume_desc* desc);
int t3d_volume_release(t3d_ctx* ctx, uint32_t volume_id);

/* Skeletons and animations for t3d_submit_poses; animation_id 0 is "no animation" (rest pose). */
int t3d_skeleton_upload(t3d_ctx* ctx, uint64_t skeleton_id, const t3d_skeleton_desc* desc);
int t3d_skeleton_release(t3d_ctx* ctx, uint64_t skeleton_id);
int t3d_animation_upload(t3d_ctx* ctx, uint64_t animation_id, const t3d_animation_desc* desc);
int t3d_animation_release(t3d_ctx* ctx, uint64_t animation_id);

/* ---- per frame ---------------------------------------------------------------------------------- */
/* Replaces: the prologue of Camera.Render that fixes vpMatrix / camera constants (camera.go:1329-1341). */
int t3d_frame_begin(t3d_ctx* ctx, const t3d_camera* cam, const t3d_world* world);
/* Drops a frame that was begun but will not be executed (the caller hit an error while submitting: a draw that needs a host
 * callback, a missing mesh): nothing runs, the previous frame's results stay valid until the next t3d_frame_begin.  A no-op
 * when no frame is open.  Replaces: nothing in the reference (its render loop cannot fail half way). */
int t3d_frame_abort(t3d_ctx* ctx);
/* Replaces: the autosubdivisionLevels argument of ProcessVertices (camera.go:1547-1553; scene.library.autosubdivisionLevels,
 * library.go:206-209): n {DistanceSquared, MaximumTriangleSize} pairs, n <= 8.  Without it (or with n = 0) meshes uploaded
 * with subdivision data draw their base triangles only. */
int t3d_frame_set_autosubdivision(t3d_ctx* ctx, const float* levels2, uint32_t n);
/* Replaces: Model.BakeLighting (model.go:1203-1268).  A frame marked this way runs ProcessVertices' transform and the light
 * loop for every submitted draw and stops there (no cull, no lists): submit the model's MeshParts as T3D_DRAW_LIT draws with
 * the light list {lights..., World.AmbientLight} (light.go beginRender / beginModel are derived on the device as always) and
 * read the sums back with t3d_frame_read_vertex_lights.  What differs from a rendered frame, as in the reference: the
 * arithmetic is always the reference's (t3d_set_exact_lighting is implied); point lights are not skipped by the per-model
 * range test (camera.go:1625-1629 is not on this path); and a light's contribution is added once per triangle corner that
 * refers to the vertex (model.go:1254-1258 walks triangles, not vertices), i.e. valence times, light by light.  The caller
 * writes RGB into VertexColors[targetChannel] (1, 1, 1 for Shadeless parts, model.go:1236-1243).  Meshes with auto-subdivision
 * data are refused (T3D_E_UNSUPPORTED). */
int t3d_frame_set_bake_lighting(t3d_ctx* ctx, int enabled);
/* The per-vertex light sums of draw `draw` of the executed frame (mesh.vertexLights; after a bake frame: the baked colours) for
 * mesh vertices [vtx_start, vtx_end) of its part: 4 floats per vertex, R G B 1.  Zeros for an unlit draw. */
int t3d_frame_read_vertex_lights(t3d_ctx* ctx, uint32_t draw, uint32_t vtx_start, uint32_t vtx_end, float* light4);
/* Replaces: Model.BakeAO, its pass over the model's own triangles (model.go:1040-1124).  The mesh must be resident
 * (t3d_mesh_upload).  tri_normals / tri_max_span: Triangle.Normal and Triangle.MaxSpan (mesh.go:2018-2037) of every triangle;
 * target_parts: one byte per MeshPart, non-zero = in AOBakeOptions.TargetMeshParts (NULL: every part, as with an empty list);
 * target_vertices: one byte per vertex for AOBakeOptions.TargetVertices (NULL: every vertex).  colors4: the n_verts x RGBA of
 * VertexColors[TargetChannel], updated in place exactly as the reference updates it (a corner that shares its position with a
 * close, sufficiently bent neighbour is pulled to occlusion_color; a vertex is pulled once per corner that refers to it, in
 * triangle order).  Synchronous.  The inter-model pass (AOBakeOptions.OtherModels, model.go:1128-1196) is not offered: it
 * indexes VertexColors by vertex where every other line indexes it by channel, so it cannot run in the reference either. */
int t3d_bake_ao(t3d_ctx* ctx, uint64_t mesh_id, const float* tri_normals3, const float* tri_max_span, const uint8_t* target_parts,
                const uint8_t* target_vertices, float occlusion_angle, const float occlusion_color[4], float* colors4);
/* Replaces: building sceneLights / LightGroup.Lights (camera.go:1275-1303, 1576-1581).  Returns a
 * light-list handle in *list_id; list 0 is always the empty list. */
int t3d_submit_lights(t3d_ctx* ctx, const t3d_light* lights, uint32_t n, uint32_t* list_id);
/* Replaces: LightVolume.beginRender / beginModel over useLightsRenderSet (light.go:565-580) and the nested
 * light.VertexLight calls of LightVolume.VertexLight (light.go:652-654).  `lights` are the lights the volume's cells
 * refer to (cell_light_index), in world space like t3d_submit_lights; ambient, point and directional only.  Per frame,
 * before the first t3d_submit_lights whose list holds the volume. */
int t3d_volume_use_lights(t3d_ctx* ctx, uint32_t volume_id, const t3d_light* lights, uint32_t n);
/* Replaces: reading bone.boneInfluence per vertex (model.go:574-580, node.go:592-594).  Appends n
 * matrices to the frame's bone arena and returns the offset of the first in *offset. */
int t3d_submit_bones(t3d_ctx* ctx, const float* matrices16, uint32_t n, uint32_t* offset);
/* Replaces: AnimationTrack.ValueAsVector / ValueAsQuaternion for every channel (animation.go:63-164, 674-722), the
 * SetLocalPosition / Scale / Rotation calls of AnimationPlayer.forceUpdate without relative motion
 * (animation.go:961-980; blending: t3d_submit_poses_blended), Node.Transform of every bone and its boneInfluence (node.go:572-594) -- for n armatures
 * playing `animation_id` on `skeleton_id`.  The device evaluates them at t3d_frame_execute; pose i's matrices are
 * bones [*offset + i * n_bones, *offset + (i + 1) * n_bones) of the same arena t3d_submit_bones fills (the
 * returned offset carries a tag bit: pass it through t3d_draw.bone_offset unchanged, adding i * n_bones). */
int t3d_submit_poses(t3d_ctx* ctx, uint64_t skeleton_id, uint64_t animation_id, const t3d_pose* poses, uint32_t n, uint32_t* offset);
/* Replaces: the same for players in the middle of a blend (AnimationPlayer.BlendTime > 0 after Play(), animation.go:548-554,
 * 852-907).  What the reference blends from is prevAnimatedProperties -- the copy Play() took of the targets of the last
 * forceUpdate under the previous animation -- which is `from_animation_id` sampled at t3d_pose.from_playhead (0: the previous
 * state animated nothing); what it blends into is `animation_id` at t3d_pose.playhead; t3d_pose.blend is the elapsed fraction.
 * Per bone and property: both exist -> start + (target - start) * blend, rotations start.Slerp(target, blend).Unit(); one
 * exists -> that one; neither -> the skeleton's rest value.  The frame in which blend reaches 1 still goes through here
 * (animation.go:899-902 ends the blend after using it); later frames use t3d_submit_poses.  A blend that starts from an
 * unfinished blend starts from a blend result: t3d_submit_pose_chains. */
int t3d_submit_poses_blended(t3d_ctx* ctx, uint64_t skeleton_id, uint64_t from_animation_id, uint64_t animation_id, const t3d_pose* poses,
                             uint32_t n, uint32_t* offset);
/* Replaces: the same when the blend was itself started in the middle of an unfinished blend, any number of times (animation.go:
 * 905-908: prevAnimatedProperties then holds blend RESULTS).  The history since the last finished blend is a chain of animations
 * animation_ids[0 .. n_levels), in the order they were Play()ed, the last one playing; the device folds it the way the reference's
 * two maps do: per bone, level i's targets are blend(targets of level i - 1, samples of level i, blend[i]) by the per-property rules
 * of t3d_submit_poses_blended where level i - 1's animation has a channel for the bone (currentProperties is rebuilt from the
 * playing animation's channels after every Play(), animation.go:618-626), the plain samples where it has not; a property no later
 * level sets keeps the value the earlier level gave the bone's Node.  "Has a channel" is taken to mean that one of the bone's three
 * tracks holds keyframes.  n_levels = 1 is t3d_submit_poses, 2 is t3d_submit_poses_blended; up to T3D_POSE_CHAIN_MAX. */
int t3d_submit_pose_chains(t3d_ctx* ctx, uint64_t skeleton_id, const uint64_t* animation_ids, uint32_t n_levels, const t3d_pose_chain* poses,
                           uint32_t n, uint32_t* offset);
/* Replaces: one render(rp) call (camera.go:1507-1827) per element, in render order. */
int t3d_submit_draws(t3d_ctx* ctx, const t3d_draw* draws, uint32_t n);
/* Replaces: n render(rp) calls (camera.go:2083-2100: the children of a DynamicBatch owner; any run of Models sharing mesh,
 * part, material and parent) TOGETHER WITH the Node.Transform() of each of those Models (node.go:572-583): the device builds
 *     transform_i = Scale(scale_i) . rotation . Translate(position_i) [. parent]
 * with the reference's Matrix4.Mult, operation for operation, instead of receiving 16 floats per Model.  Equivalent to
 * t3d_submit_draws of n copies of *proto with that transform, color = instances[i].color x proto->color (component-wise: pass
 * Material.Color, or ones without a material, in proto->color) and T3D_DRAW_CONTINUE_FLUSH set on every copy but the first
 * (which keeps proto's flag).  proto->transform is ignored; proto->world_rotation is every instance's (they share rotation and
 * parent).  rotation16: the Node.rotation the instances share, NULL = identity; parent16: parent.Transform(), NULL = the
 * Models have no parent (node.go:581).  Both must be affine (column 3 = 0,0,0,1), as every Node.Transform() of a scene graph is.
 * frustum_bounds9: NULL, or mesh.Dimensions as {Center x y z, Min x y z, Max x y z} for Models with FrustumCulling: the device then
 * also evaluates Model.onTransformUpdate's culling sphere (model.go:142-182) and Camera.SphereInFrustum (camera.go:933-983,
 * the test of camera.go:2088-2093 / 1361) per instance, and an instance outside the frustum is not rendered: it emits nothing
 * and its t3d_draw_result reports total_tris = 0 (the reference never reaches render() for it, camera.go:1532). */
int t3d_submit_instances(t3d_ctx* ctx, const t3d_draw* proto, const float* rotation16, const float* parent16, const float* frustum_bounds9,
                         const t3d_instance* instances, uint32_t n);
/* Replaces: Model.ProcessVertices + the lighting and emit loops for all submitted draws.
 * Asynchronous: returns after enqueueing on the context's stream. */
int t3d_frame_execute(t3d_ctx* ctx);
/* Waits for the frame and returns host-visible (pinned) result lists: what flush() consumes. */
int t3d_frame_results(t3d_ctx* ctx, t3d_frame_output* out, t3d_timings* timings);
/* Waits for the frame and returns DEVICE pointers (no D2H of the lists; `draws` is still host memory).
 * For device-resident consumers (NCCL gather, interop rasterisers). */
int t3d_frame_results_device(t3d_ctx* ctx, t3d_frame_output* out, t3d_timings* timings);
/* With this on, t3d_frame_results of a frame whose camera has T3D_CAM_RENDER_DEPTH off does not copy the depth list to the host
 * and returns depth_list = NULL: Camera.flush only draws it when RenderDepth is on (camera.go:1880-1934), and without RenderDepth
 * the reference's depthVertexList holds positions and the colours of an earlier frame (camera.go:1681-1708 vs 1744-1782).  Halves
 * the bytes such a frame moves over PCIe.  Off by default (both lists always come back). */
int t3d_set_skip_unused_depth_list(t3d_ctx* ctx, int enabled);
/* Two frames in flight.  With this on, t3d_frame_results still waits for the frame's kernels and returns counts, per-draw
 * results and list pointers, but only STARTS the device -> host copy of the lists (on a stream of its own, into one of two sets
 * of pinned buffers; the kernels alternate between two sets of device lists likewise): the caller builds, submits and executes
 * the next frame while the copy runs, and reads frame k's lists after t3d_frame_download_wait.  The pointers of frame k (lists
 * and `draws`) stay valid until t3d_frame_results of frame k + 2; `index_list` is one table for all frames, an identity that only
 * ever grows (it may move when a later frame has a larger flush: take it from the latest output).  This is the order the reference's game loop allows: Camera.flush's
 * DrawTriangles calls (camera.go:1829-2048) of frame k are issued while Camera.Render of frame k + 1 runs.
 * Ignored (the lists are complete on return, as without it) for frames written to an output target or delivered to another rank. */
int t3d_set_async_download(t3d_ctx* ctx, int enabled);
/* Blocks until the lists of the last frame whose results were requested (frames_back = 0) or of the one before (1) have arrived. */
int t3d_frame_download_wait(t3d_ctx* ctx, uint32_t frames_back);

/* ---- output placement ------------------------------------------------------------------------------ */
/* Makes the emit kernel write colorVertexList / depthVertexList / normalVertexList (tetra3d.go:21-23) into
 * caller-provided DEVICE memory instead of the library's own buffers: another GPU's memory mapped over
 * NVLink (object-wise sharding: every rank emits straight into the lists of the rank that issues the draws,
 * camera.go:1829-2048 consumes them there), or a buffer shared with the rasteriser.  `vertex_capacity` must
 * cover the frame's worst case, 3 vertices per submitted triangle.  normal may be NULL unless RENDER_NORMALS.
 * t3d_frame_results_device then returns these pointers; t3d_frame_results still copies to pinned host memory.
 * Passing color = NULL restores the library's buffers. */
int t3d_set_output_target(t3d_ctx* ctx, t3d_vertex* color, t3d_vertex* depth, t3d_vertex* normal, uint64_t vertex_capacity);
/* Device memory that other processes of the same machine can map (CUDA IPC): the owning rank allocates, sends
 * the handle over any channel, the other ranks open it and pass the pointer (plus their offset) to
 * t3d_set_output_target.  Freed / unmapped by t3d_destroy. */
typedef struct t3d_shared_handle { unsigned char bytes[64]; } t3d_shared_handle;
int t3d_shared_alloc(t3d_ctx* ctx, uint64_t bytes, void** device_ptr, t3d_shared_handle* handle);
int t3d_shared_open(t3d_ctx* ctx, const t3d_shared_handle* handle, void** device_ptr);

/* ---- delivery of sharded frames --------------------------------------------------------------------- */
/* Object-wise sharding: every rank runs the geometry stage on the flushes it owns (each render(rp) is
 * independent: own MVP, cull, min/max, bin sort -- model.go:602); the lists have to end up on the ONE rank that issues
 * the DrawTriangles calls (camera.go:1829-2048).  These entry points do that delivery inside the frame, device to
 * device over NVLink, with no host round trip and no collective:
 *   - a shard's emit kernel stores a 36-byte compact form of every vertex ({DstX DstY SrcX SrcY | R G B A}, the depth
 *     value, Custom0 under perspective-corrected texturing) straight into its inbox in the issuing rank's memory, ships its
 *     per-draw statistics and posts {frame serial, vertex count} to its mailbox there;
 *   - the issuing rank's frame ends with one kernel per shard that waits for that shard's post, rebuilds its part of
 *     colorVertexList / depthVertexList (tetra3d.go:21-23) and acknowledges, which is what lets the shard overwrite its
 *     inbox with the next frame.
 * Every rank calls t3d_frame_execute once per frame, in lockstep (the serial is the frame count since attach).
 *   t3d_gather_alloc   issuing rank: allocates the lists (world regions of vertex_capacity vertices), the inboxes and the
 *                      mailboxes; returns three handles for the other ranks (send them over any channel).
 *   t3d_gather_attach  every rank (the issuing rank passes handles = NULL): from the next frame on, frames are delivered.
 *                      vertex_capacity >= 3 x the triangles any rank submits in a frame, draw_capacity >= its draws.
 *   t3d_gather_results issuing rank, after t3d_frame_execute: waits for the frame; rank r's vertices are
 *                      color_list / depth_list [r * vertex_capacity, r * vertex_capacity + n_vertices[r * n_vertices_stride]).
 *   t3d_gather_rank_stats  issuing rank: rank r's per-draw {min depth bits, max depth bits, survivors} in r's submit order --
 *                      with the (deterministic) assignment of flushes to ranks this gives every flush's slice.
 * A shard that does not report within 10 s makes t3d_gather_results fail with T3D_E_CUDA instead of hanging the GPU.
 * Replaces: nothing in the reference (it is single-process); SURVEY.md 8e's "NCCL gather of draw buffers". */
typedef struct t3d_gather_output {
    uint32_t world, n_vertices_stride;
    uint64_t vertex_capacity;
    const t3d_vertex* color_list;       /* DEVICE memory on the issuing rank */
    const t3d_vertex* depth_list;
    const uint32_t* n_vertices;         /* host memory: n_vertices[r * n_vertices_stride] */
} t3d_gather_output;
typedef struct t3d_draw_stat { uint32_t min_depth_bits, max_depth_bits, survivors, reserved; } t3d_draw_stat;
int t3d_gather_alloc(t3d_ctx* ctx, uint32_t world, uint64_t vertex_capacity, uint32_t draw_capacity, t3d_shared_handle handles[3]);
int t3d_gather_attach(t3d_ctx* ctx, uint32_t world, uint32_t rank, uint32_t dst, uint64_t vertex_capacity, uint32_t draw_capacity, const t3d_shared_handle handles[3]);
int t3d_gather_detach(t3d_ctx* ctx);
int t3d_gather_results(t3d_ctx* ctx, t3d_gather_output* out);
int t3d_gather_rank_stats(t3d_ctx* ctx, uint32_t rank, uint32_t n_draws, t3d_draw_stat* out);

/* ---- lighting precision ---------------------------------------------------------------------------- */
/* Per-vertex lighting (light.go:191-258, 332-360) runs by default with single-instruction reciprocals and fused
 * multiply-adds: lit colours match the reference within 1/255 (BASELINE.json north_star; measured maximum in
 * profiles/README.md), everything else -- positions, the post-cull triangle set, depths, bins, order, depth colours --
 * stays bit-exact, including the two cut-off DECISIONS of a point light (range, 1/d*energy < 0.001).  enabled = 1
 * switches to the reference's arithmetic operation by operation (IEEE division and square root, no contraction):
 * colours bit-exact, the fused vertex kernel about 2x slower.  Not callable while a frame is open (T3D_E_STATE).
 * Replaces: nothing in the reference (it has one precision); this is the knob SURVEY.md A.8 allows. */
int t3d_set_exact_lighting(t3d_ctx* ctx, int enabled);

/* ---- introspection (tests, profiling) ------------------------------------------------------------ */
/* Stage timings put a CUDA event between the kernels of a frame, which serialises them; by default
 * (enabled = 0) the kernels are chained with programmatic dependent launches and only total_ms is measured.
 * Replaces: the per-phase timers of DebugInfo (camera.go:41-48). */
int t3d_set_stage_timings(t3d_ctx* ctx, int enabled);
/* Copies the per-vertex intermediates of draw `draw` (globalVertexTransforms tetra3d.go:30 and
 * mesh.vertexLights) for mesh vertices [vtx_start, vtx_end) of its part into clip4 / light4
 * (4 floats per vertex each; either may be NULL). */
int t3d_debug_read_draw_vertices(t3d_ctx* ctx, uint32_t draw, uint32_t vtx_start, uint32_t vtx_end,
                                 float* clip4, float* light4);
/* Model transforms and bone matrices whose column 3 is exactly (0, 0, 0, 1) cross PCIe as 12 floats.  enabled = 1
 * ships all 16 always (what happens by itself as soon as one matrix of a frame is not affine); for tests. */
/* Reads n matrices of the executed frame's bone arena back; `offset` as it would be passed in t3d_draw.bone_offset
 * (so offsets returned by t3d_submit_poses read the device-evaluated poses). */
int t3d_debug_read_bones(t3d_ctx* ctx, uint32_t offset, uint32_t n, float* matrices16);
int t3d_debug_force_full_upload(t3d_ctx* ctx, int enabled);
/* Host-side helper of the default lighting, exported for tests (needs no device): the largest squared distance d at which
 * PointLight.VertexLight's energy test `1 / d * energy < 0.001` (light.go:205-210) still lets an unranged light through. */
float t3d_debug_energy_cutoff(float energy);
/* Number of kernels launched by the last t3d_frame_execute. */
int t3d_debug_last_launch_count(t3d_ctx* ctx);
/* Which large draws (9 .. 64 tiles of 1024 triangles) are culled and sorted by the single persistent kernel instead of the
 * three-kernel segmented sort: 0 none, 1 (default) when all their tiles are resident at once, 2 always.  Both paths produce the
 * same lists bit for bit; this is a performance heuristic made switchable for tests and measurements. */
int t3d_debug_set_fused_sort(t3d_ctx* ctx, int mode);
/* Profiling builds (-DT3D_VERTEX_TRACE) only, T3D_E_UNSUPPORTED otherwise: 16 nanosecond time stamps per block of the last
 * k_vertex launch (t3d_kernels.cu lists them), for the first n_blocks <= 1024 blocks. */
int t3d_debug_vertex_trace(t3d_ctx* ctx, unsigned long long* out, uint32_t n_blocks);
/* The CUDA stream (cudaStream_t) the context launches on, for event timing by the caller. */
void* t3d_debug_stream(t3d_ctx* ctx);
/* Algorithmic bytes of the fused vertex kernel for the last frame (DESIGN.md roofline figure). */
uint64_t t3d_debug_vertex_kernel_bytes(t3d_ctx* ctx);
/* Bytes copied host->device by the last t3d_frame_execute and device->host by the last t3d_frame_results*. */
int t3d_debug_frame_bytes(t3d_ctx* ctx, uint64_t* h2d_bytes, uint64_t* d2h_bytes);

#ifdef __cplusplus
}
#endif
#endif /* T3DGEO_H */

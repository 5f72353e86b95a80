//! Drop-in for `crates/bvh` (same public names and signatures: crates/bvh/src/lib.rs:5-7) that forwards to the
//! sm_100a CUDA library through the C ABI of include/bvh_cuda.h.  Swap `bvh = { path = "crates/bvh" }` for
//! `bvh = { package = "bvh_cuda", path = "rust/bvh_cuda" }` in crates/pools/Cargo.toml and the root Cargo.toml.
//!
//! NOT compiled in this repository's image (no cargo/rustc); kept in sync with the header by hand.
use bytemuck::{Pod, Zeroable};
use components::{Instance, MeshInfo};
use glam::{UVec3, UVec4, Vec3, Vec4};
use std::os::raw::{c_char, c_int, c_void};

#[repr(C)]
#[derive(Copy, Clone, Default, Debug, Pod, Zeroable)]
pub struct BvhNode {
    pub min: Vec3,
    pub left_first: u32,
    pub max: Vec3,
    pub count: u32,
}

#[repr(C)]
#[derive(Copy, Clone, Default, Debug, Pod, Zeroable)]
pub struct TlasNode {
    pub min: Vec3,
    pub left_right: u32,
    pub max: Vec3,
    pub instance_idx: u32,
}

impl TlasNode {
    pub fn is_leaf(&self) -> bool {
        self.left_right == 0
    }
}

#[derive(PartialOrd, PartialEq, Clone, Copy, Debug)]
pub enum Dist {
    Hit(f32),
    Miss,
}

#[derive(Clone, Copy, Default, Debug)]
pub struct Ray {
    pub orig: Vec3,
    pub dir: Vec3,
}

impl Ray {
    pub fn new(orig: Vec3, dir: Vec3) -> Self {
        Self { orig, dir }
    }
}

#[allow(non_camel_case_types)]
type ctx_t = c_void;
#[allow(non_camel_case_types)]
type scene_t = c_void;

/// BvhCudaSceneDesc of include/bvh_cuda.h: the six storage buffers of voidin's trace bind group
/// (crates/pools/src/mesh/mod.rs:136-238) plus the optional unpacked TLAS child pairs.
#[repr(C)]
struct SceneDesc {
    tlas_nodes: *const TlasNode, n_tlas_nodes: usize,
    tlas_children: *const u32,
    instances: *const Instance, n_instances: usize,
    meshes: *const MeshInfo, n_meshes: usize,
    bvh_nodes: *const BvhNode, n_bvh_nodes: usize,
    vertices: *const f32, n_vertices: usize,
    indices: *const u32, n_indices: usize,
}

extern "C" {
    fn bvh_cuda_create(device: c_int, out: *mut *mut ctx_t) -> c_int;
    fn bvh_cuda_last_error(ctx: *const ctx_t) -> *const c_char;
    fn bvh_cuda_blas_build(
        ctx: *mut ctx_t, vertices: *const f32, n_vertices: usize, indices: *mut u32, n_tris: usize,
        nodes_out: *mut BvhNode, nodes_cap: usize, n_nodes_out: *mut u32,
    ) -> c_int;
    fn bvh_cuda_tlas_build(
        ctx: *mut ctx_t, instances: *const Instance, n_inst: usize, meshes: *const MeshInfo, n_mesh: usize,
        nodes_out: *mut TlasNode, children_out: *mut u32,
    ) -> c_int;
    fn bvh_cuda_trace_blas(
        ctx: *mut ctx_t, nodes: *const BvhNode, n_nodes: usize, vertices: *const f32, n_vertices: usize,
        indices: *const u32, n_tris: usize, ray_o: *const f32, ray_d: *const f32, n_rays: usize,
        t_out: *mut f32, tri_out: *mut u32,
    ) -> c_int;
    fn bvh_cuda_trace_blas_recursive(
        ctx: *mut ctx_t, nodes: *const BvhNode, n_nodes: usize, vertices: *const f32, n_vertices: usize,
        indices: *const u32, n_tris: usize, ray_o: *const f32, ray_d: *const f32, n_rays: usize,
        node_idx: u32, t0: f32, t_out: *mut f32, hit_out: *mut u8,
    ) -> c_int;
    fn bvh_cuda_blas_build_batch_dev(
        ctx: *mut ctx_t, d_vertices: *const f32, n_vertices: usize, d_indices: *mut u32, n_indices: usize,
        d_mesh_info: *mut MeshInfo, n_meshes: usize, d_nodes_out: *mut BvhNode, nodes_cap: usize,
        n_nodes_out: *mut u32, stream: *mut c_void,
    ) -> c_int;
    fn bvh_cuda_blas_build_batch_async_dev(
        ctx: *mut ctx_t, d_vertices: *const f32, n_vertices: usize, d_indices: *mut u32, n_indices: usize,
        d_mesh_info: *mut MeshInfo, n_meshes: usize, d_nodes_out: *mut BvhNode, nodes_cap: usize,
        d_result: *mut u32, stream: *mut c_void,
    ) -> c_int;
    fn bvh_cuda_blas_build_finish(ctx: *mut ctx_t, n_nodes_out: *mut u32) -> c_int;
    fn bvh_cuda_scene_upload(ctx: *mut ctx_t, host_desc: *const SceneDesc, out: *mut *mut scene_t) -> c_int;
    fn bvh_cuda_scene_instance_boxes_dev(ctx: *mut ctx_t, scene: *mut scene_t, enable: c_int, stream: *mut c_void) -> c_int;
    fn bvh_cuda_scene_free(ctx: *mut ctx_t, scene: *mut scene_t);
    fn bvh_cuda_trace_closest(
        ctx: *mut ctx_t, scene: *const scene_t, ray_o: *const f32, ray_d: *const f32, n_rays: usize, tmax: f32,
        t_out: *mut f32, tri_out: *mut u32, inst_out: *mut u32,
    ) -> c_int;
    fn bvh_cuda_trace_any(
        ctx: *mut ctx_t, scene: *const scene_t, ray_o: *const f32, ray_d: *const f32, n_rays: usize, tmax: f32,
        occluded_out: *mut u8,
    ) -> c_int;
}

thread_local! {
    // one context per host thread (the reference path is single-threaded: crates/app/src/lib.rs:110-114)
    static CTX: *mut ctx_t = unsafe {
        let mut p: *mut ctx_t = std::ptr::null_mut();
        let rc = bvh_cuda_create(0, &mut p);
        assert!(rc == 0, "bvh_cuda_create failed ({rc}): a CUDA device is required, there is no CPU fallback");
        p
    };
}

fn check(ctx: *mut ctx_t, rc: c_int) {
    if rc != 0 {
        let msg = unsafe { std::ffi::CStr::from_ptr(bvh_cuda_last_error(ctx)) };
        // the reference panics on the same inputs (blas.rs:84, mesh/mod.rs:321) or never returns (blas.rs:115,139)
        panic!("bvh_cuda error {rc}: {}", msg.to_string_lossy());
    }
}

pub struct BvhBuilder<'a> {
    num_bins: usize,
    vertices: &'a [Vec3],
    indices: &'a mut [UVec3],
}

impl<'a> BvhBuilder<'a> {
    pub fn new(vertices: &'a [Vec3], indices: &'a mut [UVec3]) -> Self {
        Self { num_bins: 8, vertices, indices }
    }

    /// Stored and never read, exactly like crates/bvh/src/blas.rs:64-67 / :136.
    pub fn set_bin_number(mut self, num_bins: usize) -> Self {
        self.num_bins = num_bins;
        self
    }

    pub fn build(self) -> Bvh {
        let n = self.indices.len();
        let mut nodes = vec![BvhNode::default(); n * 2];
        let mut used = 0u32;
        CTX.with(|&ctx| unsafe {
            let rc = bvh_cuda_blas_build(
                ctx, self.vertices.as_ptr() as *const f32, self.vertices.len(),
                self.indices.as_mut_ptr() as *mut u32, n, nodes.as_mut_ptr(), nodes.len(), &mut used,
            );
            check(ctx, rc);
        });
        nodes.truncate(used as usize);
        Bvh { nodes }
    }
}

pub struct Bvh {
    pub nodes: Vec<BvhNode>,
}

impl Bvh {
    pub fn traverse_iter(&self, vertices: &[Vec3], indices: &[UVec3], ray: Ray) -> Dist {
        let (mut t, mut tri) = (0f32, 0u32);
        CTX.with(|&ctx| unsafe {
            let rc = bvh_cuda_trace_blas(
                ctx, self.nodes.as_ptr(), self.nodes.len(), vertices.as_ptr() as *const f32, vertices.len(),
                indices.as_ptr() as *const u32, indices.len(), &ray.orig as *const Vec3 as *const f32,
                &ray.dir as *const Vec3 as *const f32, 1, &mut t, &mut tri,
            );
            check(ctx, rc);
        });
        if tri == u32::MAX { Dist::Miss } else { Dist::Hit(t) }
    }

    /// `Bvh::traverse` (crates/bvh/src/blas.rs:211-245), the recursive variant: Vec4 / UVec4 slices whose w lanes
    /// are ignored (`truncate()`), start node and distance bound supplied by the caller, `Hit(t)` as soon as the
    /// start node's box is hit (t stays at the caller's value when no triangle is closer).
    pub fn traverse(&self, vertices: &[Vec4], indices: &[UVec4], ray: Ray, node_idx: usize, t: f32) -> Dist {
        let v3: Vec<Vec3> = vertices.iter().map(|v| v.truncate()).collect();
        let i3: Vec<UVec3> = indices.iter().map(|i| i.truncate()).collect();
        let (mut t_out, mut hit) = (0f32, 0u8);
        CTX.with(|&ctx| unsafe {
            let rc = bvh_cuda_trace_blas_recursive(
                ctx, self.nodes.as_ptr(), self.nodes.len(), v3.as_ptr() as *const f32, v3.len(),
                i3.as_ptr() as *const u32, i3.len(), &ray.orig as *const Vec3 as *const f32,
                &ray.dir as *const Vec3 as *const f32, 1, node_idx as u32, t, &mut t_out, &mut hit,
            );
            check(ctx, rc);
        });
        if hit != 0 { Dist::Hit(t_out) } else { Dist::Miss }
    }
}

/// MeshPool::add for a whole pooled scene in one call (crates/pools/src/mesh/mod.rs:309-351) on DEVICE buffers
/// (e.g. wgpu/Vulkan buffers imported into CUDA): every mesh's BLAS is built in the same passes.  `mesh_info[m]`
/// carries vertex_offset / base_index / index_count on entry and bvh_index on return.  Returns the total node count.
///
/// # Safety
/// All pointers are device pointers valid for the stated sizes; `stream` is a `cudaStream_t` (null = default stream).
pub unsafe fn build_pooled_scene_dev(
    d_vertices: *const f32, n_vertices: usize, d_indices: *mut u32, n_indices: usize, d_mesh_info: *mut MeshInfo,
    n_meshes: usize, d_nodes_out: *mut BvhNode, nodes_cap: usize, stream: *mut c_void,
) -> u32 {
    let mut used = 0u32;
    CTX.with(|&ctx| {
        let rc = bvh_cuda_blas_build_batch_dev(
            ctx, d_vertices, n_vertices, d_indices, n_indices, d_mesh_info, n_meshes, d_nodes_out, nodes_cap,
            &mut used, stream,
        );
        check(ctx, rc);
    });
    used
}

/// The same forest build, stream-ordered: everything is enqueued on `stream` and the call returns at once; `d_result`
/// (device, 4 x u32, may be null) receives {nodes, status bits, interior nodes, 0}.  `finish_pooled_scene_build` waits,
/// checks the status (panics like the blocking build) and returns the node count.  One build in flight per thread context.
///
/// # Safety
/// As `build_pooled_scene_dev`; the buffers must stay valid until `finish_pooled_scene_build` returns.
pub unsafe fn enqueue_pooled_scene_build_dev(
    d_vertices: *const f32, n_vertices: usize, d_indices: *mut u32, n_indices: usize, d_mesh_info: *mut MeshInfo,
    n_meshes: usize, d_nodes_out: *mut BvhNode, nodes_cap: usize, d_result: *mut u32, stream: *mut c_void,
) {
    CTX.with(|&ctx| {
        let rc = bvh_cuda_blas_build_batch_async_dev(
            ctx, d_vertices, n_vertices, d_indices, n_indices, d_mesh_info, n_meshes, d_nodes_out, nodes_cap, d_result, stream,
        );
        check(ctx, rc);
    });
}

pub fn finish_pooled_scene_build() -> u32 {
    let mut used = 0u32;
    CTX.with(|&ctx| unsafe { check(ctx, bvh_cuda_blas_build_finish(ctx, &mut used)) });
    used
}

/// Result of `Scene::traverse_tlas`: the CUDA twin of the WGSL `TraceResult` (shaders/utils/bvh.wgsl:18-24), with the
/// ids the shader does not return.
#[derive(Clone, Copy, Debug)]
pub struct TraceResult {
    pub hit: bool,
    pub dist: f32,
    pub triangle: u32,
    pub instance: u32,
}

/// The trace bind group (crates/pools/src/mesh/mod.rs:136-238) uploaded once; `traverse_tlas` / `occluded` then
/// have the semantics of shaders/utils/bvh.wgsl:89-123 and src/bin/raytraced_shadows.wgsl:98-102.
pub struct Scene {
    handle: *mut scene_t,
}

impl Scene {
    #[allow(clippy::too_many_arguments)]
    pub fn upload(
        tlas: &Tlas, instances: &[Instance], meshes: &[MeshInfo], bvh_nodes: &[BvhNode], vertices: &[Vec3],
        indices: &[u32],
    ) -> Self {
        let desc = SceneDesc {
            tlas_nodes: tlas.nodes.as_ptr(), n_tlas_nodes: tlas.nodes.len(),
            tlas_children: if tlas.children.is_empty() { std::ptr::null() } else { tlas.children.as_ptr() as *const u32 },
            instances: instances.as_ptr(), n_instances: instances.len(),
            meshes: meshes.as_ptr(), n_meshes: meshes.len(),
            bvh_nodes: bvh_nodes.as_ptr(), n_bvh_nodes: bvh_nodes.len(),
            vertices: vertices.as_ptr() as *const f32, n_vertices: vertices.len(),
            indices: indices.as_ptr(), n_indices: indices.len(),
        };
        let mut handle: *mut scene_t = std::ptr::null_mut();
        CTX.with(|&ctx| unsafe { check(ctx, bvh_cuda_scene_upload(ctx, &desc, &mut handle)) });
        Self { handle }
    }

    /// `traverse_tlas(ray)` for a batch of rays (closest hit).
    pub fn traverse_tlas(&self, rays: &[Ray]) -> Vec<TraceResult> {
        let n = rays.len();
        let o: Vec<Vec3> = rays.iter().map(|r| r.orig).collect();
        let d: Vec<Vec3> = rays.iter().map(|r| r.dir).collect();
        let (mut t, mut tri, mut inst) = (vec![0f32; n], vec![0u32; n], vec![0u32; n]);
        CTX.with(|&ctx| unsafe {
            check(ctx, bvh_cuda_trace_closest(
                ctx, self.handle, o.as_ptr() as *const f32, d.as_ptr() as *const f32, n, 1e30,
                t.as_mut_ptr(), tri.as_mut_ptr(), inst.as_mut_ptr(),
            ))
        });
        (0..n).map(|i| TraceResult { hit: tri[i] != u32::MAX, dist: t[i], triangle: tri[i], instance: inst[i] }).collect()
    }

    /// Shadow rays: `traverse_tlas(ray).hit` with early exit (src/bin/raytraced_shadows.wgsl:98-102).
    pub fn occluded(&self, rays: &[Ray]) -> Vec<bool> {
        let n = rays.len();
        let o: Vec<Vec3> = rays.iter().map(|r| r.orig).collect();
        let d: Vec<Vec3> = rays.iter().map(|r| r.dir).collect();
        let mut occ = vec![0u8; n];
        CTX.with(|&ctx| unsafe {
            check(ctx, bvh_cuda_trace_any(ctx, self.handle, o.as_ptr() as *const f32, d.as_ptr() as *const f32, n, 1e30, occ.as_mut_ptr()))
        });
        occ.into_iter().map(|b| b != 0).collect()
    }
}

impl Scene {
    /// Recompute the tight per-instance world boxes the exact-order kernels cull with (uploaded scenes have them from the
    /// start; only scenes wrapped around caller-owned device buffers need this after the instance buffer was rewritten).
    pub fn refresh_instance_boxes(&self) {
        CTX.with(|&ctx| unsafe { check(ctx, bvh_cuda_scene_instance_boxes_dev(ctx, self.handle, 1, std::ptr::null_mut())) });
    }
}

impl Drop for Scene {
    fn drop(&mut self) {
        CTX.with(|&ctx| unsafe { bvh_cuda_scene_free(ctx, self.handle) });
    }
}

pub struct Tlas {
    pub nodes: Vec<TlasNode>,
    /// unpacked child ids per node (needed by a traversal above 32 767 instances; left_right packs 16+16 bits)
    pub children: Vec<[u32; 2]>,
}

impl Tlas {
    pub fn empty() -> Self {
        Self { nodes: vec![], children: vec![] }
    }

    pub fn build(&mut self, instances: &[Instance], meshes: &[MeshInfo]) {
        let total = 2 * instances.len() + 1;
        self.nodes = vec![TlasNode::default(); total];
        self.children = vec![[0u32; 2]; total];
        CTX.with(|&ctx| unsafe {
            let rc = bvh_cuda_tlas_build(
                ctx, instances.as_ptr(), instances.len(), meshes.as_ptr(), meshes.len(),
                self.nodes.as_mut_ptr(), self.children.as_mut_ptr() as *mut u32,
            );
            check(ctx, rc);
        });
    }
}

// Drop-in replacement for the body of upstream `src/mc_code.rs`: same public
// signature, the history loop runs on a B200 through the C ABI of
// include/nraps_mc.h.  No CPU fallback: a missing device is a panic, like every
// other error in the upstream function.
use crate::{DeltaX, Mesh, SolutionResults, Variables, XSData};
use std::os::raw::{c_char, c_int};

#[repr(C)]
struct NrapsProblem {
    m: u32, g: u32, n: u32, nf: u32, numass: u32,
    generations: u64, histories: u64, skip: u64,
    boundl: f32, boundr: f32, dx_fuel: f32, dx_water: f32, k0: f32,
    sigt: *const f32, sigs: *const f32, mu: *const f32, siga: *const f32, sigf: *const f32,
    nut: *const f32, chit: *const f32, inv_sigtr: *const f32, scat: *const f32,
    matid: *const u8, dx: *const f32, left: *const f32, right: *const f32,
    fuel_indices: *const u64,
}

#[repr(C)]
#[derive(Default)]
struct NrapsOptions {
    seed: u64, stream: u64, stride: u64,
    device: i32, scatter_mode: i32, stale_xs: i32, source_mode: i32, tracking_mode: i32,
    kernel_variant: i32, threads_per_block: i32, blocks_per_sm: i32, chunk: i32, quiet: i32,
    bank_cap: i32, spawn_batch: i32, walk_cap: i32, slots_per_thread: i32,
    max_flights: u64,
    profile_phases: i32, reserved0: i32,
}

#[repr(C)]
struct NrapsResults {
    flux: *mut f32, assembly_average: *mut f32, fission_source: *mut f32, k: *mut f32, k_fund: *mut f32,
    tally_fixed: *mut u64, counters: [u64; 8], seconds_device: f64,
    bank_sizes: *mut u64, entropy: *mut f64, flux_moments: *mut f64,
}

extern "C" {
    fn nraps_mc_run(p: *const NrapsProblem, o: *const NrapsOptions, r: *mut NrapsResults) -> c_int;
    fn nraps_strerror(code: c_int) -> *const c_char;
}

pub fn monte_carlo(
    variables: &Variables,
    xsdata: &XSData,
    delta_x: &DeltaX,
    meshid: &Vec<Mesh>,
    fuel_indices: &Vec<usize>,
    k_new: f32,
) -> SolutionResults {
    let n = meshid.len();
    let g = variables.energygroups as usize;
    // Vec<Mesh> (AoS) -> structure of arrays
    let matid: Vec<u8> = meshid.iter().map(|c| c.matid).collect();
    let dx: Vec<f32> = meshid.iter().map(|c| c.delta_x).collect();
    let left: Vec<f32> = meshid.iter().map(|c| c.mesh_left).collect();
    let right: Vec<f32> = meshid.iter().map(|c| c.mesh_right).collect();
    let fuel: Vec<u64> = fuel_indices.iter().map(|&i| i as u64).collect();

    let mut flux = vec![0f32; g * n];
    let mut avg = vec![0f32; g * n];
    let mut fission = vec![0f32; n];
    let mut k = vec![0f32; variables.generations];
    let mut k_fund = vec![0f32; variables.generations];

    let p = NrapsProblem {
        m: variables.mattypes as u32, g: g as u32, n: n as u32, nf: fuel.len() as u32, numass: variables.numass as u32,
        generations: variables.generations as u64, histories: variables.histories as u64, skip: variables.skip as u64,
        boundl: variables.boundl, boundr: variables.boundr, dx_fuel: delta_x.fuel, dx_water: delta_x.water, k0: k_new,
        sigt: xsdata.sigt.as_ptr(), sigs: xsdata.sigs.as_ptr(), mu: xsdata.mu.as_ptr(), siga: xsdata.siga.as_ptr(),
        sigf: xsdata.sigf.as_ptr(), nut: xsdata.nut.as_ptr(), chit: xsdata.chit.as_ptr(),
        inv_sigtr: xsdata.inv_sigtr.as_ptr(), scat: xsdata.scat_matrix.as_ptr(),
        matid: matid.as_ptr(), dx: dx.as_ptr(), left: left.as_ptr(), right: right.as_ptr(), fuel_indices: fuel.as_ptr(),
    };
    let o = NrapsOptions { stale_xs: 1, ..Default::default() };
    let mut r = NrapsResults {
        flux: flux.as_mut_ptr(), assembly_average: avg.as_mut_ptr(), fission_source: fission.as_mut_ptr(),
        k: k.as_mut_ptr(), k_fund: k_fund.as_mut_ptr(), tally_fixed: std::ptr::null_mut(), counters: [0; 8],
        seconds_device: 0.0, bank_sizes: std::ptr::null_mut(), entropy: std::ptr::null_mut(), flux_moments: std::ptr::null_mut(),
    };
    let rc = unsafe { nraps_mc_run(&p, &o, &mut r) };
    if rc != 0 {
        let msg = unsafe { std::ffi::CStr::from_ptr(nraps_strerror(rc)) };
        panic!("nraps_mc_run failed: {}", msg.to_string_lossy());
    }
    SolutionResults {
        flux: flux.chunks(n).map(|row| row.to_vec()).collect(),
        assembly_average: avg.chunks(n).map(|row| row.to_vec()).collect(),
        fission_source: fission,
        k,
        k_fund,
    }
}

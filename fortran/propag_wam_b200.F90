! propag_wam_b200.F90 -- PROPAG_WAM with the reference's call signature (src/ecwam/propag_wam.F90:10-11, dummies as :74-77), body
! on the B200: replaces src/ecwam/propag_wam.F90 in an ecWAM build that links libecwam_b200.so.
!
! What the reference body does and where it went: FL1 -> FL1_EXT block copy (:105-147) and the copy back (:368-405): gone, the
! kernel gathers from / writes the NPROMA-chunked arrays; MPEXCHNG (:166) -> NCCL send/recv inside the call; first-call CTUWUPDT +
! PROENVHALO (:221-236) and PROPDOT (:171-216) -> the library's CTU set-up, redone after ECWAM_B200_INVALIDATE_WEIGHTS (LUPDTWGHT /
! LLUPDTTD, getcurr.F90:285-289); PROPAGS2 incl. the fast-wave sub-steps (:239-313) -> propags2 kernels.  IPROPAGS = 2, structured
! grid only (LLUNSTR, IPROPAGS = 0, 1 are rejected at ECWAM_B200_CREATE).

SUBROUTINE PROPAG_WAM (BLK2GLO, WAVNUM, CGROUP, OMOSNH2KD, FL1, &
&  DEPTH, DELLAM1, COSPHM1, UCUR, VCUR)

      USE PARKIND_WAVE, ONLY : JWIM, JWRB
      USE YOWDRVTYPE  , ONLY : WVGRIDGLO
      USE YOWGRID  , ONLY : NPROMA_WAM, NCHNK
      USE YOWPARAM , ONLY : NANG     ,NFRE
      USE YOWUBUF  , ONLY : LUPDTWGHT
      USE YOWREFD  , ONLY : LLUPDTTD
      USE YOWABORT , ONLY : WAM_ABORT
      USE YOMHOOK  , ONLY : LHOOK,   DR_HOOK, JPHOOK
      USE ECWAM_B200_MOD, ONLY : B200_HANDLE, ECWAM_B200_PROPAG_WAM_F, ECWAM_B200_INVALIDATE_WEIGHTS, ECWAM_B200_ERRMSG
      USE, INTRINSIC :: ISO_C_BINDING, ONLY : C_INT

      IMPLICIT NONE

      TYPE(WVGRIDGLO), INTENT(IN) :: BLK2GLO
      REAL(KIND=JWRB), DIMENSION(NPROMA_WAM, NANG, NFRE, NCHNK), INTENT(INOUT) :: FL1
      REAL(KIND=JWRB), DIMENSION(NPROMA_WAM, NFRE, NCHNK), INTENT(IN) :: WAVNUM, CGROUP, OMOSNH2KD
      REAL(KIND=JWRB), DIMENSION(NPROMA_WAM, NCHNK), INTENT(IN) :: DEPTH, DELLAM1, COSPHM1, UCUR, VCUR

      INTEGER(KIND=C_INT) :: IERR
      CHARACTER(LEN=16) :: CLNUM
      REAL(KIND=JPHOOK) :: ZHOOK_HANDLE

! ----------------------------------------------------------------------

      IF (LHOOK) CALL DR_HOOK('PROPAG_WAM',0,ZHOOK_HANDLE)

!     New currents / refraction terms: the CTU weights have to be rebuilt (propag_wam.F90:171-236)
      IF (LUPDTWGHT .OR. LLUPDTTD) THEN
        IERR = ECWAM_B200_INVALIDATE_WEIGHTS(B200_HANDLE)
        LUPDTWGHT = .FALSE.
        LLUPDTTD = .FALSE.
      ENDIF

!$acc host_data use_device(WAVNUM, CGROUP, OMOSNH2KD, FL1, DEPTH, DELLAM1, COSPHM1, UCUR, VCUR)
      IERR = ECWAM_B200_PROPAG_WAM_F(B200_HANDLE, WAVNUM, CGROUP, OMOSNH2KD, FL1, DEPTH, DELLAM1, COSPHM1, UCUR, VCUR)
!$acc end host_data

!     > 0: number of grid points that violate the CFL / weight-range checks of CTUW (ctuwdrv.F90:127-146 aborts there)
      IF (IERR > 0) THEN
        WRITE(CLNUM,'(I16)') IERR
        CALL WAM_ABORT('PROPAG_WAM (ecwam_b200): CFL VIOLATED AT '//TRIM(ADJUSTL(CLNUM))//' GRID POINTS',__FILENAME__,__LINE__)
      ELSEIF (IERR < 0) THEN
        CALL WAM_ABORT('PROPAG_WAM (ecwam_b200): '//TRIM(ECWAM_B200_ERRMSG()),__FILENAME__,__LINE__)
      ENDIF

      IF (LHOOK) CALL DR_HOOK('PROPAG_WAM',1,ZHOOK_HANDLE)

END SUBROUTINE PROPAG_WAM
